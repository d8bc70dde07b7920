/*
 * gsk_oracle.h — interface of the CPU oracle (TEST INFRASTRUCTURE; see gsk_oracle.c header).
 * Takes the same gsk_problem as libgskrige.so so that tests hand identical bytes to both.
 */
#ifndef GSK_ORACLE_H
#define GSK_ORACLE_H
#include "../include/gskrige.h"
#ifdef __cplusplus
extern "C" {
#endif
enum { GSK_ORACLE_SEARCH_BRUTE = 0, GSK_ORACLE_SEARCH_KDTREE = 1 };
/* restates exactsolve / approxsolve (ref: src/estimation/krig.jl:166-234) */
int gsk_oracle_krige(const gsk_problem *p, double *mean_out, double *var_out, int32_t *nneigh_out,
                     int32_t *neigh_idx_out, int search_kind, int nthreads);
/* restates search! (ref: src/estimation/krig.jl:210) only */
int gsk_oracle_search(const gsk_problem *p, int32_t *nneigh_out, int32_t *neigh_idx_out, double *d2_out,
                      int search_kind, int nthreads);
/* restates the simulation loop of SeqSim with SGS's estimator and marginal (ref: src/simulation/seq.jl:102-135,
 * src/simulation/sgs.jl:62-69); argument meaning as gsk_sgs_plan + gsk_sgs_sample of include/gskrige.h */
int gsk_oracle_sgs(int dim, int64_t n, const double *const *coords, const int64_t *rank, int vario_kind,
                   double vario_range, double vario_sill, double vario_nugget, double gaussian_nugget_eps, double mean,
                   int min_neighbors, int max_neighbors, double ball_radius, const double *values, const double *z,
                   double *out, int32_t *nneigh_out, int32_t *neigh_idx_out, double *weights_out, double *sigma_out);
int gsk_oracle_uk_exponents(int degree, int dim, int32_t *out, int cap);
int64_t gsk_oracle_num_targets(const gsk_problem *p);
int gsk_oracle_threads(void);
/* drops the searcher (packed coordinates + KD-tree) kept from the last local gsk_oracle_krige call */
void gsk_oracle_clear_cache(void);
#ifdef __cplusplus
}
#endif
#endif
