// map.cuh -- apop_map with a materialised result (see map.cu)
#pragma once
#include "common.cuh"
#include <stddef.h>

namespace apb {
// cell functions (values match apop_b200_map_fn 0..5 in include/apop_b200.h, then the additions)
enum MapCellFn { MAPC_IDENTITY = 0, MAPC_DEV = 1, MAPC_DEV2 = 2, MAPC_POISSON = 3, MAPC_LOG = 4, MAPC_EXP = 5,
                 MAPC_ABS = 6, MAPC_SQRT = 7, MAPC_SCALE = 8, MAPC_POW = 9, MAPC_COUNT = 10 };
// row functions (apop_b200_row_fn)
enum MapRowFn { MAPR_SUM = 0, MAPR_MEAN = 1, MAPR_SUMSQ = 2, MAPR_MAX = 3, MAPR_MIN = 4, MAPR_DOT = 5,
                MAPR_PROBIT_LL = 6, MAPR_LOGIT2_LL = 7, MAPR_POISSON_REG_LL = 8, MAPR_OLS_ERROR = 9, MAPR_COUNT = 10 };
cudaError_t map_cells_launch(const double* tiles, size_t n_tiles, int k, int ycol, int cols, int fn, double p0, bool do_matrix,
                             bool do_vector, double* out_tiles, cudaStream_t st);
cudaError_t map_rows_launch(const double* tiles, size_t row0, size_t rows, size_t n_rows_total, int k, int ycol, int cols, int fn,
                            const double* param, const double* aux, double* out, cudaStream_t st);
}  // namespace apb
