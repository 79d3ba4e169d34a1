/* api_smoke.c -- the reference's API smoke test (runners/api/src/lib.rs:102-126) in plain C against
 * the two public headers: three coincident particles, five ticks, capacity-sized read-back.
 *
 *   gcc -std=c11 -Wall -Wextra -pedantic -Iinclude examples/api_smoke.c \
 *       -Lwrach_b200/lib -lwrach_cuda -Wl,-rpath,$PWD/wrach_b200/lib -o /tmp/api_smoke && /tmp/api_smoke
 *
 * Exit status: 0 on a B200; 3 with the library's message when there is no CUDA device (there is no
 * CPU fallback); 1 on any other failure.  tests/test_abi.py builds and runs it. */
#include <stdio.h>

#include "wrach_cuda.h"
#include "wrach_host.h"

int main(void) {
    wrach_config config;
    wrach_config_default(&config);         /* config_app.rs:24-34: 480 x 352, cell 3 */
    config.dimensions[0] = 10;             /* lib.rs:105-108 */
    config.dimensions[1] = 10;
    wrach_api *api = NULL;
    int rc = wrach_api_new(&config, 0, WRACH_ARITH_SPV, &api);
    if (rc != WRACH_OK) {
        fprintf(stderr, "wrach_api_new: status %d: %s\n", rc, wrach_api_last_error(NULL));
        return rc == WRACH_ERR_CUDA ? 3 : 1;
    }
    const float particles[3][4] = {{1.f, 1.f, 0.1f, 0.1f}, {1.f, 1.f, 0.1f, 0.1f}, {1.f, 1.f, 0.1f, 0.1f}};
    if (wrach_api_add_particles(api, &particles[0][0], 3) != WRACH_OK) return 1;
    for (int t = 0; t < 5; t++)
        if (wrach_api_tick(api) != WRACH_OK) {
            fprintf(stderr, "tick %d: %s\n", t, wrach_api_last_error(api));
            return 1;
        }
    uint64_t n_pos = 0, n_vel = 0;
    const float *pos = wrach_api_positions(api, &n_pos), *vel = wrach_api_velocities(api, &n_vel);
    printf("read back %llu positions, %llu velocities; particle 0 at (%g, %g) moving (%g, %g)\n",
           (unsigned long long)n_pos, (unsigned long long)n_vel, pos[0], pos[1], vel[0], vel[1]);
    /* lib.rs:122-125: the whole buffer comes back (164 slots for a 10 x 10 world), and it moved */
    const int ok = n_pos == 164 && n_vel == 164 && !(pos[0] == 0.f && pos[1] == 0.f) && !(vel[0] == 0.f && vel[1] == 0.f);
    wrach_api_free(api);
    return ok ? 0 : 1;
}
