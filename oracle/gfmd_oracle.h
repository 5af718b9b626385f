/* ORACLE / TEST INFRASTRUCTURE ONLY -- see gfmd_oracle.c */
#ifndef GFMD_ORACLE_H
#define GFMD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif
int gfmd_oracle_gather(int nall, int nlocal, const double *x, const double *xeq,
                       int *gid, const int *mask, int groupbit,
                       int nx, int ny, int xlo_loc, int ylo_loc, int nx_loc, int ny_loc,
                       double xprd, double yprd, int dxshift, int dyshift,
                       double *u_xy);
double gfmd_oracle_post_force(int nx, int ny, int ndof, const double *phi,
                              const double *linf, const double *u, double *f,
                              double *u0, int fft_backend);
int gfmd_oracle_scatter(int nall, int nlocal, const int *gid, const int *mask,
                        int groupbit, int xlo_loc, int xhi_loc, int ylo_loc,
                        int yhi_loc, const double *f_xy, double *f,
                        double *fsum_loc);
#ifdef __cplusplus
}
#endif
#endif
