/* Plain C99 consumer of include/cpl_batched.h: the header must be valid C and the host-only entry points must work
 * without a GPU (tests/test_abi_host.py builds and runs this with gcc -std=c99 -pedantic -Wall -Werror). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cpl_batched.h"

int main(void)
{
    const char *names[4] = {"r_foot", "l_foot", "r_hand", "l_hand"};
    cplb_problem *p = NULL;
    int32_t n, m, nnz, row, col;
    double lb[3], ub[3], w;
    if (cplb_abi_version() != CPLB_ABI_VERSION) return 1;
    if (cplb_create(4, names, CPLB_ENV_GROUND, 100.0, -1, &p) != CPLB_OK) { puts(cplb_last_error()); return 2; }
    if (cplb_get_dims(p, &n, &m, &nnz) != CPLB_OK || n != 39 || m != 30 || nnz != 174) return 3;
    {
        int32_t *iRow = (int32_t *)malloc(sizeof(int32_t) * (size_t)nnz), *jCol = (int32_t *)malloc(sizeof(int32_t) * (size_t)nnz);
        uint8_t *is_const = (uint8_t *)malloc((size_t)nnz);
        int k, nconst = 0;
        if (cplb_get_jacobian_structure(p, iRow, jCol) != CPLB_OK) return 4;
        if (iRow[0] != 0 || jCol[0] != 3 || iRow[nnz - 1] != m - 1) return 5;
        if (cplb_get_jacobian_constants(p, is_const, NULL) != CPLB_OK) return 6;
        for (k = 0; k < nnz; k++) nconst += is_const[k];
        if (nconst != 72) return 7;
        free(iRow); free(jCol); free(is_const);
    }
    if (cplb_get_contact_row(p, "l_foot", &row) != CPLB_OK || row != 6) return 8;          /* sorted rank 0 */
    if (cplb_get_block_column(p, CPLB_BLOCK_POSITION, "l_foot", &col) != CPLB_OK || col != 15) return 9; /* vector index 1 */
    if (cplb_set_mu(p, -1.0) != CPLB_INVALID_ARGUMENT || strstr(cplb_last_error(), "friction") == NULL) return 10;
    if (cplb_set_force_threshold(p, "nose", 1.0) != CPLB_OUT_OF_RANGE) return 11;
    lb[0] = lb[1] = lb[2] = 0.0; ub[0] = ub[1] = 1.0; ub[2] = -1.0;
    if (cplb_set_bounds(p, CPLB_BLOCK_FORCE, "r_hand", lb, ub) != CPLB_INVALID_ARGUMENT) return 12;
    if (cplb_set_com_weight(p, 2.0) != CPLB_OK || cplb_get_com_weight(p, &w) != CPLB_OK || w != 2.0) return 13;
    {
        cplb_eval_args a;
        cplb_instance_params ip;
        memset(&a, 0, sizeof a);
        memset(&ip, 0, sizeof ip);
        a.num_instances = 0;             /* an empty batch is a successful no-op on any machine */
        a.per_instance = &ip;
        if (cplb_eval_host(p, &a) != CPLB_OK) return 14;
    }
    cplb_destroy(p);
    puts("C ABI ok");
    return 0;
}
