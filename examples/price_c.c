/* price_c.c -- the C-ABI from plain C99 (no CUDA headers, no C++): what a cgo / JNI / ctypes binding sees.
 *   gcc -std=c99 -Iinclude examples/price_c.c -Lmonte-carlo-project-cuda_b200 -lmcb200 -lm -o build/price_c
 *
 *   price_c            one engine on device 0
 *   price_c 0,1,2,3    ONE engine over four GPUs (mcb_engine_create_multi): same calls, same bits
 *   price_c 0,0        two shards on one GPU (how the single-GPU tests exercise the sharded path) */
#define _POSIX_C_SOURCE 199309L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mcb200.h"

int main(int argc, char **argv)
{
    int devices[MCB_MAX_PEERS] = {0};
    int n_devices = 1;
    if (argc > 1) {
        n_devices = 0;
        char *list = argv[1];
        for (char *tok = strtok(list, ","); tok && n_devices < MCB_MAX_PEERS; tok = strtok(NULL, ","))
            devices[n_devices++] = atoi(tok);
    }
    mcb_engine *engine = NULL;
    if (mcb_engine_create_multi(devices, n_devices, &engine) != MCB_OK) {
        fprintf(stderr, "no engine: %s\n", mcb_last_error());
        return 2;   /* no CPU fallback: without a B200 this is the expected outcome */
    }
    mcb_option_data opt = {100.0f, 1.0f, 100.0f, 0.05f, 0.2f, 120.0f, 10, 50, 1 << 20, 1000, 100, 0.01f};
    mcb_result call, put, bullet;
    int rc = mcb_price_european(engine, &opt, 0, 1234, MCB_CALL, &call);
    rc |= mcb_price_european(engine, &opt, 0, 1234, MCB_PUT, &put);
    rc |= mcb_price_bullet(engine, &opt, 0, 1234, 0, 0.0f, 0, &bullet);
    float *rows = (float *)malloc(sizeof(float) * 8 * 100);
    rc |= mcb_simulate_trajectories(engine, &opt, 0, 8, 1234, rows, NULL, MCB_HOST);
    /* invalid input comes back as a status code and a message, never as an exit */
    mcb_option_data bad = opt;
    bad.S0 = -1.0f;
    const int bad_rc = mcb_price_european(engine, &bad, 0, 1234, MCB_CALL, &call);
    printf("C_ABI %d %d %.17g %.17g %.17g %.17g %.9g %.9g\n", rc, bad_rc, call.price, call.std_error, put.price, bullet.price,
           (double)rows[0], (double)rows[799]);
    printf("parity C - P = %.6f, S0 - K e^{-rT} = %.6f\n", call.price - put.price, 100.0 - 100.0 * exp(-0.05));

    /* pipelined: several jobs in flight (one launch per shard each), collected afterwards; a 2^27-path
     * job last, so the line below also says how long a large job takes through the synchronous call */
    uint64_t tickets[MCB_PIPELINE_DEPTH];
    mcb_result piped[MCB_PIPELINE_DEPTH];
    for (int i = 0; i < MCB_PIPELINE_DEPTH; ++i)
        rc |= mcb_european_submit(engine, &opt, (uint64_t)(100000 + 1000003 * i), 1234, MCB_CALL, &tickets[i]);
    for (int i = 0; i < MCB_PIPELINE_DEPTH; ++i) rc |= mcb_european_collect(engine, tickets[i], &piped[i]);
    printf("C_PIPE %d %d", rc, mcb_engine_shard_count(engine));
    for (int i = 0; i < MCB_PIPELINE_DEPTH; ++i) printf(" %.17g %.17g", piped[i].sum, piped[i].sumsq);
    printf("\n");

    float strikes[3] = {90.0f, 100.0f, 110.0f}, vols[3] = {0.15f, 0.2f, 0.3f};
    mcb_result sweep[3];
    rc |= mcb_price_sweep(engine, &opt, strikes, vols, 3, 300000, 1234, MCB_CALL, sweep);
    mcb_option_data nm = opt;
    nm.N_STEPS = 12;
    nm.step = 1.0f / 12.0f;
    nm.N_PATHS_INNER = 128;
    float F[10 * 12];
    double mean_F = 0.0;
    rc |= mcb_nested_monte_carlo(engine, &nm, 0, 10, 1234, 1235, MCB_DISCOUNT_CORRECT, F, NULL, NULL, MCB_HOST, &mean_F);
    printf("C_MORE %d %.17g %.17g %.17g %.9g %.9g %.17g\n", rc, sweep[0].sum, sweep[1].sum, sweep[2].sum, (double)F[0],
           (double)F[119], mean_F);
    /* latency of the synchronous call, from the reference's own job size (hello.cu:14: 1e5 paths) down to
     * one path (pure call overhead: launch + ticket + final tree + host-visible result) */
    {
        const uint64_t sizes[6] = {1, 16384, 100000, 1000000, 16384, 1};   /* (the small ones twice: first and last) */
        for (int k = 0; k < 6; ++k) {
            struct timespec t0, t1;
            mcb_result r;
            const int reps = 2000;
            for (int i = 0; i < (k ? 500 : 5000); ++i)   /* (the first size also lets the clocks settle) */
                rc |= mcb_price_european(engine, &opt, sizes[k], 1234, MCB_CALL, &r);
            clock_gettime(CLOCK_MONOTONIC, &t0);
            for (int i = 0; i < reps; ++i) rc |= mcb_price_european(engine, &opt, sizes[k], 1234, MCB_CALL, &r);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            const double us = ((double)(t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec)) / 1e3 / reps;
            printf("C_LAT %d shards, mcb_price_european(%llu paths): %.2f us per synchronous call\n",
                   mcb_engine_shard_count(engine), (unsigned long long)sizes[k], us);
        }
    }
    free(rows);
    mcb_engine_destroy(engine);
    return rc;
}
