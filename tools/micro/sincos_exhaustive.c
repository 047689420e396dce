#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <omp.h>
static void sc32(float a, float* s, float* c) {
    float kf = rintf(a * 0.63661977236758134308f);
    float r = fmaf(-kf, 1.5703125f, a);
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    float z = r * r;
    float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float sr = fmaf(sp * z, r, r);
    float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    float cr = fmaf(cp, z * z, fmaf(-0.5f, z, 1.0f));
    int k = ((int)kf) & 3;
    *s = (k == 0) ? sr : (k == 1) ? cr : (k == 2) ? -sr : -cr;
    *c = (k == 0) ? cr : (k == 1) ? -sr : (k == 2) ? -cr : sr;
}
int main() {
    float hi = 6.2831855f; uint32_t uh; memcpy(&uh, &hi, 4);
    double max_abs = 0, max_ulp = 0; uint64_t n = 0, off1 = 0;
    #pragma omp parallel
    {
        double ma = 0, mu = 0; uint64_t cnt = 0, o1 = 0;
        #pragma omp for schedule(static)
        for (int64_t i = 0; i <= (int64_t)uh + 16; i++) {
            uint32_t u = (uint32_t)i; float a; memcpy(&a, &u, 4);
            float s, c; sc32(a, &s, &c);
            double rs = sin((double)a), rc = cos((double)a);
            float fs = (float)rs, fc = (float)rc;
            double es = fabs((double)s - rs), ec = fabs((double)c - rc);
            if (es > ma) ma = es; if (ec > ma) ma = ec;
            // error in units of the ulp of the reference value (1.0 scale floor: ulp(1) = 1.19e-7 is what matters for directions)
            double us = es / fmax(fabs(rs) * 1.1920929e-7, 1e-45), uc = ec / fmax(fabs(rc) * 1.1920929e-7, 1e-45);
            if (fabs(rs) > 1e-3 && us > mu) mu = us; if (fabs(rc) > 1e-3 && uc > mu) mu = uc;
            if (s != fs) o1++; if (c != fc) o1++;
            cnt++;
        }
        #pragma omp critical
        { if (ma > max_abs) max_abs = ma; if (mu > max_ulp) max_ulp = mu; n += cnt; off1 += o1; }
    }
    printf("n=%llu max abs err %.3e, max rel err (|value|>1e-3) %.2f ulp, results differing from the correctly rounded value: %.2f %%\n",
           (unsigned long long)n, max_abs, max_ulp, 100.0 * off1 / (2.0 * n));
    return 0;
}
