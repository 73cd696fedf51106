/* ntl_jit.h -- NVRTC specialisation of the scan kernel (ntl_scan.cuh) for one pattern set; see ntl_jit.cpp */
#ifndef NTL_JIT_H
#define NTL_JIT_H
#include <cuda_runtime.h>
#include <string>
#include "ntl_dev.h"

struct ntl_jit_kernel;

/* Compile ntl_scan.cuh with the pattern set of *prm baked in, for sm_<major><minor>a, and load it on the current
 * device.  Returns nullptr (and the reason in *err) if NVRTC is not available or the build fails. */
ntl_jit_kernel *ntl_jit_build(const ntl_dev_params *prm, int major, int minor, std::string *err);
/* Compile only (no device needed): returns the cubin size or < 0; log receives the NVRTC log / error. */
long ntl_jit_compile(const ntl_dev_params *prm, const char *arch, std::string *cubin, std::string *log);
std::string ntl_jit_source(const ntl_dev_params *prm);      /* the generated prologue + kernel entry */
/* four_bit: the build for reads with IUPAC letters (four nibble planes) instead of the 2-bit one */
cudaError_t ntl_jit_launch(ntl_jit_kernel *k, const ntl_scan_args *a, int four_bit, int grid, cudaStream_t st);
int ntl_jit_blocks_per_sm(ntl_jit_kernel *k, int four_bit);
void ntl_jit_free(ntl_jit_kernel *k);
#endif
