"""CPU checks of arithmetic shortcuts the kernels take: each must produce the SAME BITS as the operation it replaces."""
import os
import subprocess
import tempfile
import textwrap

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fma_division_by_the_hemisphere_pdf_is_the_ieee_quotient():
    """div_by_hemisphere_pdf (nrcu_shade.cuh): q = x y, r = fma(-q, c, x), q' = fma(r, y, q) with c = 1/(2 pi), y = 1/c equals
    x / c bit for bit for 1e-30 <= |x| <= 1e30.  tools/micro/div_by_pdf_exhaustive.c walks all 2^32 patterns (28 s on 8
    cores, 0 mismatches in the range); here: every 1021st pattern plus the boundaries, through the kernel's own header."""
    src = textwrap.dedent("""
        #include <cstdio>
        #include <cstdint>
        #include <cstring>
        #include <cmath>
        #define NRCU_HOST_EMU 1
        #include "nrcu_shade.cuh"
        int main() {
            const float c = 1 / (2 * NRCU_PT_PI);
            unsigned long long bad = 0, n = 0;
            for (uint64_t i = 0; i < (1ull << 32); i += 1021) {
                uint32_t u = (uint32_t)i; float x; std::memcpy(&x, &u, 4);
                if (x != x) continue;
                volatile float ref = x / c;
                float got = nrcu::div_by_hemisphere_pdf(x), r = ref;
                uint32_t a, b; std::memcpy(&a, &r, 4); std::memcpy(&b, &got, 4);
                n++; if (a != b) bad++;
            }
            const float edge[] = {0.f, -0.f, 1e-30f, -1e-30f, 1e30f, -1e30f, 9.9999e-31f, 1.0001e30f, 1.f, 0.72499996f};
            for (float x : edge) {
                volatile float ref = x / c; float got = nrcu::div_by_hemisphere_pdf(x), r = ref;
                uint32_t a, b; std::memcpy(&a, &r, 4); std::memcpy(&b, &got, 4);
                n++; if (a != b) bad++;
            }
            std::printf("%llu %llu\\n", n, bad);
            return bad != 0;
        }
    """)
    with tempfile.TemporaryDirectory() as td:
        cpp, exe = os.path.join(td, "t.cpp"), os.path.join(td, "t")
        open(cpp, "w").write(src)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{REPO}/include", f"-I{REPO}/nrenderer_b200/csrc", cpp, "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        n, bad = map(int, r.stdout.split())
        assert r.returncode == 0 and bad == 0 and n > 4_000_000, r.stdout
