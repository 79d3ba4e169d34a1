/* pair_trace.c -- for every cell of a packed frame: n9 = min(count, 9) and a 36-bit mask of which of
 * its pairs (reference order, particles.rs:62-83) push.  Input for tools/model/pair_phase_model.py,
 * which replays warp schedules of the pair phase on real push patterns.  Same arithmetic as the
 * checker (arith = spv). Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC -o libpairtrace.so pair_trace.c -lm */
#include <math.h>
#include <stdint.h>

void pair_trace(const uint32_t *indices, const float *pos, uint32_t cells, uint8_t *n9_out, uint64_t *mask_out) {
    for (uint32_t c = 0; c < cells; c++) {
        uint32_t start = indices[c + 1], all = indices[c + 2] - start, n = all > 9 ? 9 : all;
        float p[18];
        for (uint32_t i = 0; i < n; i++) { p[2 * i] = pos[2 * (start + i)]; p[2 * i + 1] = pos[2 * (start + i) + 1]; }
        uint64_t mask = 0;
        uint32_t bit = 0;
        for (uint32_t l = 0; l < n; l++)
            for (uint32_t r = l + 1; r < n; r++, bit++) {
                float *L = p + 2 * l, *R = p + 2 * r;
                float dx = L[0] - R[0], dy = L[1] - R[1];
                float d = sqrtf(fmaf(dx, dx, dy * dy));
                if (d > 1.0f) continue;
                if (d == 0.0f) d = 0.0001f;
                float force = 0.5f * (1.0f - d) / d, vx = R[0] - L[0], vy = R[1] - L[1];
                float lx = fmaf(-vx, force, L[0]), ly = fmaf(-vy, force, L[1]);
                float rx = fmaf(vx, force, R[0]), ry = fmaf(vy, force, R[1]);
                L[0] = lx; L[1] = ly; R[0] = rx; R[1] = ry;
                mask |= 1ull << bit;
            }
        n9_out[c] = (uint8_t)n;
        mask_out[c] = mask;
    }
}
