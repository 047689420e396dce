"""Flat scenes: the POD mirror of NRenderer::Scene used across the C ABI.

`FlatScene` holds numpy arrays named exactly like the fields of `nrcu_scene` (include/nrcu.h),
which itself mirrors the reference's `Scene` (reference code/include/scene/Scene.hpp:40-67) in
model-local coordinates.  `.nrsc` files (written by nr_headless --dump-flat from the reference's
own importers, or by `FlatScene.save`) are the fixtures under tests/golden/.

This module is host plumbing only: no rendering arithmetic lives here.
"""
from __future__ import annotations

import copy
import ctypes as C
import struct
from dataclasses import dataclass, field

import numpy as np

_DTYPES = {0: np.float32, 1: np.uint32, 2: np.int32, 3: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}

MATERIAL_PARAM_FLOATS = 22
# offsets of the float parameters inside nrcu_material after (type, present)
MP = {"diffuse_color": (0, 3, 1 << 0), "specular_color": (3, 3, 1 << 1), "specular_ex": (6, 1, 1 << 2),
      "albedo": (7, 3, 1 << 3), "eta_r": (10, 3, 1 << 4), "eta_i": (13, 3, 1 << 5), "ior": (16, 1, 1 << 6),
      "absorbed": (17, 3, 1 << 7), "roughness": (20, 1, 1 << 8), "f0": (21, 1, 1 << 9)}

MODE_RAYCAST, MODE_SIMPLE, MODE_ACC = 0, 1, 2
NODE_SPHERE, NODE_TRIANGLE, NODE_PLANE, NODE_MESH = 0, 1, 2, 3
GLASS_STOCHASTIC, GLASS_BRANCH = 0, 1


class NrcuMaterial(C.Structure):
    _fields_ = [("type", C.c_uint32), ("present", C.c_uint32), ("params", C.c_float * MATERIAL_PARAM_FLOATS)]


_P = C.c_void_p


class NrcuScene(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32), ("samples_per_pixel", C.c_uint32),
        ("cam_position", C.c_float * 3), ("cam_up", C.c_float * 3), ("cam_look_at", C.c_float * 3),
        ("cam_fov", C.c_float), ("cam_aperture", C.c_float), ("cam_focus_distance", C.c_float), ("cam_aspect", C.c_float),
        ("ambient_type", C.c_uint32), ("ambient_constant", C.c_float * 3), ("ambient_environment_map", C.c_int32),
        ("n_models", C.c_uint32), ("model_translation", _P),
        ("n_nodes", C.c_uint32), ("node_type", _P), ("node_entity", _P), ("node_model", _P),
        ("n_spheres", C.c_uint32), ("sphere_position", _P), ("sphere_radius", _P), ("sphere_material", _P),
        ("n_triangles", C.c_uint32), ("triangle_vertices", _P), ("triangle_normal", _P), ("triangle_material", _P),
        ("n_planes", C.c_uint32), ("plane_normal", _P), ("plane_position", _P), ("plane_u", _P), ("plane_v", _P),
        ("plane_material", _P),
        ("n_meshes", C.c_uint32), ("mesh_vertex_offset", _P), ("mesh_index_offset", _P), ("mesh_positions", _P),
        ("mesh_indices", _P), ("mesh_material", _P),
        ("n_materials", C.c_uint32), ("materials", _P),
        ("n_point_lights", C.c_uint32), ("point_intensity", _P), ("point_position", _P),
        ("n_area_lights", C.c_uint32), ("area_radiance", _P), ("area_position", _P), ("area_u", _P), ("area_v", _P),
        ("n_textures", C.c_uint32), ("texture_width", _P), ("texture_height", _P), ("texture_offset", _P),
        ("texture_rgba", _P),
    ]


def _f32(*shape):
    return field(default_factory=lambda: np.zeros(shape, np.float32))


def _u32(*shape):
    return field(default_factory=lambda: np.zeros(shape, np.uint32))


def _i32(*shape):
    return field(default_factory=lambda: np.zeros(shape, np.int32))


@dataclass
class FlatScene:
    width: int = 500
    height: int = 500
    depth: int = 4
    samples_per_pixel: int = 16
    cam_position: np.ndarray = field(default_factory=lambda: np.array([0, 0, 10], np.float32))
    cam_up: np.ndarray = field(default_factory=lambda: np.array([0, 1, 0], np.float32))
    cam_look_at: np.ndarray = field(default_factory=lambda: np.array([0, 0, 1000], np.float32))
    cam_fov: float = 40.0
    cam_aperture: float = 0.0
    cam_focus_distance: float = 0.1
    cam_aspect: float = 1.0
    ambient_type: int = 0
    ambient_constant: np.ndarray = _f32(3)
    ambient_environment_map: int = -1
    model_translation: np.ndarray = _f32(0, 3)
    node_type: np.ndarray = _u32(0)
    node_entity: np.ndarray = _u32(0)
    node_model: np.ndarray = _u32(0)
    sphere_position: np.ndarray = _f32(0, 3)
    sphere_radius: np.ndarray = _f32(0)
    sphere_material: np.ndarray = _i32(0)
    triangle_vertices: np.ndarray = _f32(0, 9)
    triangle_normal: np.ndarray = _f32(0, 3)
    triangle_material: np.ndarray = _i32(0)
    plane_normal: np.ndarray = _f32(0, 3)
    plane_position: np.ndarray = _f32(0, 3)
    plane_u: np.ndarray = _f32(0, 3)
    plane_v: np.ndarray = _f32(0, 3)
    plane_material: np.ndarray = _i32(0)
    mesh_vertex_offset: np.ndarray = _u32(1)
    mesh_index_offset: np.ndarray = _u32(1)
    mesh_positions: np.ndarray = _f32(0, 3)
    mesh_indices: np.ndarray = _u32(0)
    mesh_material: np.ndarray = _i32(0)
    material_type_present: np.ndarray = _u32(0, 2)
    material_params: np.ndarray = _f32(0, MATERIAL_PARAM_FLOATS)
    point_intensity: np.ndarray = _f32(0, 3)
    point_position: np.ndarray = _f32(0, 3)
    area_radiance: np.ndarray = _f32(0, 3)
    area_position: np.ndarray = _f32(0, 3)
    area_u: np.ndarray = _f32(0, 3)
    area_v: np.ndarray = _f32(0, 3)
    texture_width: np.ndarray = _u32(0)
    texture_height: np.ndarray = _u32(0)
    texture_offset: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint64))
    texture_rgba: np.ndarray = _f32(0)

    _SHAPES = {"model_translation": (-1, 3), "sphere_position": (-1, 3), "triangle_vertices": (-1, 9),
               "triangle_normal": (-1, 3), "plane_normal": (-1, 3), "plane_position": (-1, 3), "plane_u": (-1, 3),
               "plane_v": (-1, 3), "mesh_positions": (-1, 3), "material_type_present": (-1, 2),
               "material_params": (-1, MATERIAL_PARAM_FLOATS), "point_intensity": (-1, 3), "point_position": (-1, 3),
               "area_radiance": (-1, 3), "area_position": (-1, 3), "area_u": (-1, 3), "area_v": (-1, 3)}
    _ARRAYS = ["model_translation", "node_type", "node_entity", "node_model", "sphere_position", "sphere_radius",
               "sphere_material", "triangle_vertices", "triangle_normal", "triangle_material", "plane_normal",
               "plane_position", "plane_u", "plane_v", "plane_material", "mesh_vertex_offset", "mesh_index_offset",
               "mesh_positions", "mesh_indices", "mesh_material", "material_type_present", "material_params",
               "point_intensity", "point_position", "area_radiance", "area_position", "area_u", "area_v",
               "texture_width", "texture_height", "texture_offset", "texture_rgba"]

    # ------------------------------------------------------------------ io
    @classmethod
    def load(cls, path) -> "FlatScene":
        with open(path, "rb") as f:
            data = f.read()
        if data[:8] != b"NRSC0001":
            raise ValueError(f"{path}: not an NRSC0001 file")
        rec, off = {}, 8
        while off < len(data):
            (nl,) = struct.unpack_from("<I", data, off); off += 4
            name = data[off:off + nl].decode(); off += nl
            dt, cnt = struct.unpack_from("<IQ", data, off); off += 12
            dtype = np.dtype(_DTYPES[dt])
            rec[name] = np.frombuffer(data, dtype, cnt, off).copy(); off += cnt * dtype.itemsize
        s = cls()
        if "render_option" in rec:
            s.width, s.height, s.depth, s.samples_per_pixel = (int(x) for x in rec["render_option"])
        if "camera" in rec:
            c = rec["camera"]
            s.cam_position, s.cam_up, s.cam_look_at = c[0:3].copy(), c[3:6].copy(), c[6:9].copy()
            s.cam_fov, s.cam_aperture, s.cam_focus_distance, s.cam_aspect = (float(x) for x in c[9:13])
        if "ambient_type" in rec:
            s.ambient_type = int(rec["ambient_type"][0])
        if "ambient_constant" in rec:
            s.ambient_constant = rec["ambient_constant"]
        if "ambient_environment_map" in rec:
            s.ambient_environment_map = int(rec["ambient_environment_map"][0])
        for name in cls._ARRAYS:
            if name in rec:
                a = rec[name]
                if name in cls._SHAPES:
                    a = a.reshape(cls._SHAPES[name])
                setattr(s, name, a)
        if s.mesh_vertex_offset.size == 0:
            s.mesh_vertex_offset = np.zeros(1, np.uint32)
        if s.mesh_index_offset.size == 0:
            s.mesh_index_offset = np.zeros(1, np.uint32)
        return s

    def save(self, path) -> None:
        def put(f, name, arr):
            arr = np.ascontiguousarray(arr)
            f.write(struct.pack("<I", len(name))); f.write(name.encode())
            f.write(struct.pack("<IQ", _DTYPE_IDS[arr.dtype], arr.size)); f.write(arr.tobytes())
        with open(path, "wb") as f:
            f.write(b"NRSC0001")
            put(f, "render_option", np.array([self.width, self.height, self.depth, self.samples_per_pixel], np.uint32))
            cam = np.concatenate([self.cam_position, self.cam_up, self.cam_look_at,
                                  [self.cam_fov, self.cam_aperture, self.cam_focus_distance, self.cam_aspect]]).astype(np.float32)
            put(f, "camera", cam)
            put(f, "ambient_type", np.array([self.ambient_type], np.uint32))
            put(f, "ambient_constant", np.asarray(self.ambient_constant, np.float32))
            put(f, "ambient_environment_map", np.array([self.ambient_environment_map], np.int32))
            for name in self._ARRAYS:
                put(f, name, getattr(self, name))

    def copy(self) -> "FlatScene":
        return copy.deepcopy(self)

    # ------------------------------------------------------------------ editing helpers (what a GUI user does by hand)
    @property
    def n_materials(self) -> int:
        return int(self.material_type_present.shape[0])

    def add_material(self, mtype: int, **props) -> int:
        """Append a material; props use the nrcu_material field names. Returns its 0-based index."""
        params = np.zeros(MATERIAL_PARAM_FLOATS, np.float32)
        present = 0
        for k, v in props.items():
            o, n, bit = MP[k]
            params[o:o + n] = np.asarray(v, np.float32).reshape(n)
            present |= bit
        self.material_type_present = np.concatenate(
            [self.material_type_present.reshape(-1, 2), np.array([[mtype, present]], np.uint32)])
        self.material_params = np.concatenate([self.material_params.reshape(-1, MATERIAL_PARAM_FLOATS), params[None]])
        return self.n_materials - 1

    def set_triangles(self, tris: np.ndarray, material: int, translation=(0, 0, 0)) -> None:
        """Replace the geometry by one model of explicit triangles (n x 3 x 3, model-local); normals from the winding."""
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 3, 3)
        n = len(tris)
        nrm = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
        self.model_translation = np.array([translation], np.float32)
        self.node_type = np.full(n, NODE_TRIANGLE, np.uint32)
        self.node_entity = np.arange(n, dtype=np.uint32)
        self.node_model = np.zeros(n, np.uint32)
        self.triangle_vertices = tris.reshape(n, 9)
        self.triangle_normal = nrm.astype(np.float32)
        self.triangle_material = np.full(n, material, np.int32)

    def add_texture(self, rgba: np.ndarray) -> int:
        rgba = np.ascontiguousarray(rgba, np.float32)
        h, w, c = rgba.shape
        assert c == 4
        self.texture_offset = np.concatenate([self.texture_offset, np.array([self.texture_rgba.size], np.uint64)])
        self.texture_width = np.concatenate([self.texture_width, np.array([w], np.uint32)])
        self.texture_height = np.concatenate([self.texture_height, np.array([h], np.uint32)])
        self.texture_rgba = np.concatenate([self.texture_rgba.reshape(-1), rgba.reshape(-1)])
        return int(self.texture_width.size - 1)

    # ------------------------------------------------------------------ C view
    def c_view(self):
        """Return (NrcuScene, keepalive) — the ctypes struct points into arrays held by keepalive."""
        keep = []

        def ptr(a, dtype):
            a = np.ascontiguousarray(a, dtype)
            keep.append(a)
            return a.ctypes.data if a.size else None

        n_mat = self.n_materials
        mats = (NrcuMaterial * max(n_mat, 1))()
        for i in range(n_mat):
            mats[i].type = int(self.material_type_present[i, 0])
            mats[i].present = int(self.material_type_present[i, 1])
            for k in range(MATERIAL_PARAM_FLOATS):
                mats[i].params[k] = float(self.material_params[i, k])
        keep.append(mats)
        s = NrcuScene()
        s.width, s.height, s.depth, s.samples_per_pixel = self.width, self.height, self.depth, self.samples_per_pixel
        for i in range(3):
            s.cam_position[i] = float(self.cam_position[i]); s.cam_up[i] = float(self.cam_up[i])
            s.cam_look_at[i] = float(self.cam_look_at[i]); s.ambient_constant[i] = float(self.ambient_constant[i])
        s.cam_fov, s.cam_aperture = self.cam_fov, self.cam_aperture
        s.cam_focus_distance, s.cam_aspect = self.cam_focus_distance, self.cam_aspect
        s.ambient_type, s.ambient_environment_map = self.ambient_type, self.ambient_environment_map
        s.n_models = len(self.model_translation); s.model_translation = ptr(self.model_translation, np.float32)
        s.n_nodes = len(self.node_type)
        s.node_type = ptr(self.node_type, np.uint32); s.node_entity = ptr(self.node_entity, np.uint32)
        s.node_model = ptr(self.node_model, np.uint32)
        s.n_spheres = len(self.sphere_radius)
        s.sphere_position = ptr(self.sphere_position, np.float32); s.sphere_radius = ptr(self.sphere_radius, np.float32)
        s.sphere_material = ptr(self.sphere_material, np.int32)
        s.n_triangles = len(self.triangle_material)
        s.triangle_vertices = ptr(self.triangle_vertices, np.float32); s.triangle_normal = ptr(self.triangle_normal, np.float32)
        s.triangle_material = ptr(self.triangle_material, np.int32)
        s.n_planes = len(self.plane_material)
        s.plane_normal = ptr(self.plane_normal, np.float32); s.plane_position = ptr(self.plane_position, np.float32)
        s.plane_u = ptr(self.plane_u, np.float32); s.plane_v = ptr(self.plane_v, np.float32)
        s.plane_material = ptr(self.plane_material, np.int32)
        s.n_meshes = len(self.mesh_material)
        s.mesh_vertex_offset = ptr(self.mesh_vertex_offset, np.uint32); s.mesh_index_offset = ptr(self.mesh_index_offset, np.uint32)
        s.mesh_positions = ptr(self.mesh_positions, np.float32); s.mesh_indices = ptr(self.mesh_indices, np.uint32)
        s.mesh_material = ptr(self.mesh_material, np.int32)
        s.n_materials = n_mat; s.materials = C.addressof(mats)
        s.n_point_lights = len(self.point_position)
        s.point_intensity = ptr(self.point_intensity, np.float32); s.point_position = ptr(self.point_position, np.float32)
        s.n_area_lights = len(self.area_position)
        s.area_radiance = ptr(self.area_radiance, np.float32); s.area_position = ptr(self.area_position, np.float32)
        s.area_u = ptr(self.area_u, np.float32); s.area_v = ptr(self.area_v, np.float32)
        s.n_textures = len(self.texture_width)
        s.texture_width = ptr(self.texture_width, np.uint32); s.texture_height = ptr(self.texture_height, np.uint32)
        s.texture_offset = ptr(self.texture_offset, np.uint64); s.texture_rgba = ptr(self.texture_rgba, np.float32)
        return s, keep
