// exhaustive check: for every float x with |x| < 32, is the 3-op sequence
//   q0 = x*rc; rem = fma(-q0, c, x); q1 = fma(rem, rc, q0)   (rc = RN(1/c), c = 2*3.1416f)
// equal to the IEEE quotient x / c ?
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
int main(void)
{
    const float c = 2.0f * 3.1416f;
    const float rc = 1.0f / c;
    long long bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
    for (uint64_t u = 0; u < (1ull << 32); u++) {
        uint32_t b = (uint32_t)u;
        float x;
        memcpy(&x, &b, 4);
        if (!(fabsf(x) < 32.0f) || fabsf(x) < 0x1p-100f) continue;
        float q0 = x * rc;
        float rem = fmaf(-q0, c, x);
        float q1 = fmaf(rem, rc, q0);
        float ref = x / c;
        uint32_t a1, a2;
        memcpy(&a1, &q1, 4);
        memcpy(&a2, &ref, 4);
        n++;
        if (a1 != a2) {
            bad++;
            if (bad < 5) printf("mismatch x=%a q1=%a ref=%a\n", x, q1, ref);
        }
    }
    printf("checked %lld floats, mismatches %lld (c=%a rc=%a)\n", n, bad, c, rc);
    return 0;
}
