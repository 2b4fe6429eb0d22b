/* Plain-C consumer of include/gsb200.h: proves the header is valid C (no C++-isms), that the shared
 * library links from C, and that argument errors are reported without a device.  With a GPU present
 * (argv[1] == "gpu") it also runs one tiny summation and one kriging evaluation against closed forms. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "gsb200.h"

#define CHECK(cond, msg)                                                   \
    do {                                                                   \
        if (!(cond)) {                                                     \
            fprintf(stderr, "FAIL %s (%s)\n", msg, gsb_last_error());      \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char **argv)
{
    double cov[2] = {0.5, 0.25}, z1[1] = {2.0}, z2[1] = {3.0}, pos[2 * 3] = {0, 1, 2, 0, 2, 4}, out[3];
    gsb_epilogue epi;
    gsb_cov_model model;
    int ndev = -1;

    CHECK(gsb_version() >= 100, "version");
    CHECK(gsb_device_count(&ndev) == GSB_OK && ndev >= 0, "device count");
    CHECK(gsb_summate(cov, z1, z2, pos, 3, 0, 1, 3, out, GSB_MEM_HOST, 0, NULL) == GSB_ERR_ARGUMENT, "dim check");
    CHECK(strstr(gsb_last_error(), "dim") != NULL, "error text");
    memset(&epi, 0, sizeof epi);
    epi.scale = 1.0;
    epi.n_add = GSB_EPI_MAX_ADD + 1;
    CHECK(gsb_summate_ex(cov, z1, z2, pos, 3, 2, 1, 3, out, &epi, GSB_MEM_HOST, 0, NULL) == GSB_ERR_ARGUMENT,
          "epilogue check");
    memset(&model, 0, sizeof model);
    model.type = 99;
    CHECK(gsb_krige_evaluate(&model, cov, cov, 1, cov, 1, 1, pos, 3, 3, 0, NULL, 3, out, out, GSB_MEM_HOST, 0,
                             NULL) == GSB_ERR_ARGUMENT, "model check");
    CHECK(gsb_set_option("nope", 1) == GSB_ERR_ARGUMENT, "option check");
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        int i;
        double kmat[1] = {1.0 / 2.0}, cond[1] = {3.0}, cpos[1] = {0.0}, x[3] = {0.0, 1.0, 2.0}, f[3], e[3];
        CHECK(ndev > 0, "a CUDA device");
        CHECK(gsb_summate(cov, z1, z2, pos, 3, 2, 1, 3, out, GSB_MEM_HOST, 0, NULL) == GSB_OK, "summate");
        for (i = 0; i < 3; ++i) {
            const double ph = 0.5 * pos[i] + 0.25 * pos[3 + i];
            CHECK(fabs(out[i] - (2.0 * cos(ph) + 3.0 * sin(ph))) < 1e-11, "summate value");
        }
        /* simple kriging with one datum at 0, Exponential(var 2, len 1.5): field = 3 * c(x) / 2 */
        model.type = GSB_COV_EXPONENTIAL;
        model.var = 2.0;
        model.len_rescaled = 1.5;
        model.sill = 2.0;
        CHECK(gsb_krige_evaluate(&model, kmat, cond, 1, cpos, 1, 1, x, 3, 3, 0, NULL, 3, f, e, GSB_MEM_HOST, 0,
                                 NULL) == GSB_OK, "krige_evaluate");
        for (i = 0; i < 3; ++i) {
            const double c = 2.0 * exp(-x[i] / 1.5);
            CHECK(fabs(f[i] - 3.0 * c / 2.0) < 1e-13 && fabs(e[i] - c * c / 2.0) < 1e-13, "krige value");
        }
        printf("gpu ok\n");
    }
    printf("abi ok\n");
    return 0;
}
