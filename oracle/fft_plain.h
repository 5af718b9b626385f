/* ORACLE / TEST INFRASTRUCTURE ONLY -- see fft_plain.c */
#ifndef ORACLE_FFT_PLAIN_H
#define ORACLE_FFT_PLAIN_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct fftp_plan fftp_plan;
fftp_plan *fftp_plan_2d(int nx, int ny);
void fftp_destroy(fftp_plan *p);
/* interleaved complex [nx][ny], in place, unnormalised; sign -1 = exp(-i q r) */
void fftp_exec_2d(const fftp_plan *p, double *data, int sign);
void fftp_dft2d_ld(int nx, int ny, double *data, int sign);
/* threads of fftp_exec_2d: n > 0 explicit, 0 = OpenMP default */
void fftp_set_threads(int n);
int fftp_get_threads(void);
#ifdef __cplusplus
}
#endif
#endif
