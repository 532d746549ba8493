/* snf_oracle.h — C interface of the CPU oracle (test infrastructure; see snf_oracle.c). */
#ifndef SNF_ORACLE_H
#define SNF_ORACLE_H
#include "../include/snfb.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct so_result so_result;
typedef struct so_view {
    uint64_t n_leads; const snfb_lead* leads;          /* bin order: (task, svtype, bin, BAM order) */
    uint32_t n_task; uint32_t _pad; const uint32_t* task_read_count; const double* task_mean_nm; const double* rec_nm; const double* task_cov_mean;
    uint64_t n_pass, soft_errors;
    uint64_t n_cand; const snfb_cand* cand; uint64_t n_cand_leads; const snfb_lead* cand_leads;
    const uint64_t* rnames; const uint32_t* rn_off;
    uint64_t n_alt; const uint8_t* alt;
} so_view;
/* stages: 1 = leads only, 2 = + clustering/candidates/coverage, 3 = + INS consensus */
so_result* so_run(const snfb_records* R, const snfb_config* cfg, int stages, int threads);
void so_get(const so_result* r, so_view* v);
void so_free(so_result* r);
uint64_t so_hash_name(const uint8_t* s, size_t n);
uint64_t so_qname_hash(const uint8_t* s, size_t n);
double so_sqrt_frac(uint64_t p_hi, uint64_t p_lo, uint64_t q);
long so_center(const long* v, long n);
double so_stdev(const long* v, long n);
double so_stdev_trim(const long* v, long n);
int so_cigar_analyze(const uint8_t* c, int n, long out[4]);
#ifdef __cplusplus
}
#endif
#endif
