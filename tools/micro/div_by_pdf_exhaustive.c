#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <omp.h>
int main() {
    const float PI_f = 3.1415926535898f;
    volatile float pdfv = 1 / (2 * PI_f);
    const float c = pdfv;
    const float y = 1.0f / c;
    printf("c=%.9g y=%.9g\n", c, y);
    uint64_t bad = 0, bad_normal = 0;
    #pragma omp parallel for reduction(+:bad,bad_normal) schedule(static)
    for (int64_t i = 0; i < (1LL << 32); i++) {
        uint32_t u = (uint32_t)i; float x; memcpy(&x, &u, 4);
        float ref = x / c;
        float q = x * y;
        float r = fmaf(-q, c, x);
        float q2 = fmaf(r, y, q);
        uint32_t a, b; memcpy(&a, &ref, 4); memcpy(&b, &q2, 4);
        if (a != b && !(ref != ref && q2 != q2)) {
            bad++;
            float ax = fabsf(x);
            if (ax >= 1e-30f && ax <= 1e30f) { bad_normal++; if (bad_normal < 5) printf("x=%a ref=%a got=%a\n", x, ref, q2); }
        }
    }
    printf("mismatches: %llu, of which |x| in [1e-30,1e30]: %llu\n", (unsigned long long)bad, (unsigned long long)bad_normal);
    return 0;
}
