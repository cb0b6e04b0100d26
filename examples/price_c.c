/* price_c.c -- the C-ABI from plain C99 (no CUDA headers, no C++): what a cgo / JNI / ctypes binding sees.
 *   gcc -std=c99 -Iinclude examples/price_c.c -Lmonte-carlo-project-cuda_b200 -lmcb200 -lm -o build/price_c */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mcb200.h"

int main(void)
{
    mcb_engine *engine = NULL;
    if (mcb_engine_create(0, &engine) != MCB_OK) {
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
    free(rows);
    mcb_engine_destroy(engine);
    return rc;
}
