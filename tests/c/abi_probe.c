/* tests/c/abi_probe.c -- a plain C99 client of include/cg_b200.h, linked against libcg_b200.so (tests/test_capi.py).
 * What a cgo / JNI / N-API binding would do: build a scheme descriptor, create a handle, upload a configuration,
 * evaluate.  On a machine without a GPU cg_create must refuse with CG_ERR_NO_DEVICE (there is no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "cg_b200.h"

int main(void) {
    cg_scheme_desc a, b;
    const char *keys[] = {"cutoff", "alpha"};
    const double vals[] = {12.0, 0.25};
    double s[4], q = 0.5;
    cg_handle *h = NULL;
    int rc;
    printf("%s\n", cg_version());
    if (cg_scheme_wolf(12.0, 0.25, &a) != CG_OK) return 1;
    if (cg_scheme_create("wolf", 2, keys, vals, &b) != CG_OK) return 2;
    if (memcmp(&a, &b, sizeof a) != 0) return 3;
    if (cg_scheme_create("nonsense", 0, NULL, NULL, &b) != CG_ERR_INVALID) return 4;
    if (cg_srf_host(&a, 1, &q, s) != CG_OK) return 5;
    /* Wolf: S(q) = erfc(eta q) - q erfc(eta), eta = alpha * cutoff (reference wolf.hpp:60) */
    if (fabs(s[0] - (erfc(3.0 * q) - q * erfc(3.0))) > 1e-15) return 6;
    rc = cg_create(&a, -1, &h);
    if (cg_device_count() == 0) {
        if (rc != CG_ERR_NO_DEVICE || h != NULL) return 7;
        printf("no device: %s\n", cg_last_error(NULL));
        return 0;
    }
    if (rc != CG_OK) return 8;
    {
        const double x[3] = {1.0, 2.0, 9.0}, y[3] = {1.0, 1.5, 1.0}, z[3] = {1.0, 1.0, 1.0}, ch[3] = {1.0, -1.0, 1.0};
        const double box[3] = {40.0, 40.0, 40.0};
        double force[9], energy = 0.0;
        int64_t pairs = 0;
        if (cg_set_system(h, 3, x, y, z, ch, NULL, NULL, NULL, box, 1) != CG_OK) return 9;
        if (cg_eval(h, CG_MODE_ION, CG_ENERGY | CG_FORCE, force, NULL, NULL, &energy, &pairs) != CG_OK) return 10;
        if (pairs != 6) return 11; /* all three are within 12 of each other: 6 ordered pairs */
        if (fabs(force[0] + force[3] + force[6]) > 1e-12) return 12; /* pair forces cancel */
        printf("U = %.15g, pairs = %lld\n", energy, (long long)pairs);
    }
    cg_destroy(h);
    return 0;
}
