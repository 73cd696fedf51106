/* ntl_jit.h -- the compile-time-specialised scan kernel (ntl_scan.cuh) for one pattern set and span geometry:
 * precompiled cubins, the per-user cubin cache and NVRTC; see ntl_jit.cpp */
#ifndef NTL_JIT_H
#define NTL_JIT_H
#include <cuda_runtime.h>
#include <string>
#include "ntl_dev.h"

struct ntl_jit_kernel;

#define NTL_JIT_ORIGIN_PRECOMPILED 1   /* <library dir>/precompiled/ (built by build.py)  */
#define NTL_JIT_ORIGIN_CACHE       2   /* the per-user cubin cache                        */
#define NTL_JIT_ORIGIN_NVRTC       3   /* compiled now with NVRTC                         */

/* Get the specialised kernels for *prm on the current device (sm_<major><minor>a): precompiled, cached, or compiled
 * with NVRTC.  Returns nullptr (and the reason in *err) if none of the three works. */
ntl_jit_kernel *ntl_jit_build(const ntl_dev_params *prm, int major, int minor, std::string *err);
/* Compile only (no device needed): returns the cubin size or < 0; log receives the NVRTC log / error. */
long ntl_jit_compile(const ntl_dev_params *prm, const char *arch, std::string *cubin, std::string *log);
/* Compile and store <dir>/<key>.cubin in the format ntl_jit_build looks for (build.py: the precompiled directory). */
long ntl_jit_precompile(const ntl_dev_params *prm, const char *arch, const char *dir, std::string *log);
std::string ntl_jit_source(const ntl_dev_params *prm);      /* the generated prologue + kernel entries */
/* warps per CTA and dynamic shared memory of the 2-bit / 4-bit kernel for this geometry */
void ntl_jit_launch_shape(const ntl_dev_params *prm, int four_bit, int *warps, int *smem_bytes);
/* four_bit: the build for reads with IUPAC letters (four nibble planes) instead of the 2-bit one */
cudaError_t ntl_jit_launch(ntl_jit_kernel *k, const ntl_scan_args *a, int four_bit, int n_sms, cudaStream_t st);
int ntl_jit_blocks_per_sm(ntl_jit_kernel *k, int four_bit);
int ntl_jit_origin(ntl_jit_kernel *k);
void ntl_jit_free(ntl_jit_kernel *k);
#endif
