// common.cuh - runtime plumbing shared by every kernel file of librte_rrtmgp_b200.so
//
// * one launch stream PER HOST THREAD (cudaStreamPerThread until the thread calls rrtmgpb_set_stream, e.g. with
//   torch's current stream): the entry points are re-entrant across host threads (runtime.cu)
// * CUDA errors abort: the reference kernels are `subroutine`s with no error path
//   (SURVEY.md section 8b "Errors"), so a failed launch must never be silently ignored
// * DevIn/DevOut/DevInOut: pointer provenance at the ABI.  The Fortran frontend hands the
//   extern kernels whatever it allocated; device/managed pointers are used in place
//   (stream-ordered, no copies), pageable/pinned host pointers are staged through the device
//   (correctness path: copy in, run, copy out, synchronise).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cfloat>
#include <cstdint>
#include "rte_types.h"

namespace rrtmgpb {

#define RB_CUDA_CHECK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "rte_rrtmgp_b200: CUDA error %s at %s:%d: %s\n",                    \
                   cudaGetErrorName(err__), __FILE__, __LINE__, cudaGetErrorString(err__));   \
      std::abort();                                                                           \
    }                                                                                         \
  } while (0)

cudaStream_t stream();
void count_launch(int n = 1);
// Scratch in device memory (stream-ordered pool).
void* dev_alloc(size_t bytes);
void dev_free(void* p);
bool is_device_ptr(const void* p);

// Optional per-kernel timing with CUDA events on the launch stream (rrtmgpb_profile_enable()):
// construct one right before a launch; the destructor records the closing event.
struct KernelTimer {
  explicit KernelTimer(const char* name);
  ~KernelTimer();
  KernelTimer(const KernelTimer&) = delete;
  KernelTimer& operator=(const KernelTimer&) = delete;
  int slot;
  void* b_;  // the closing event (cudaEvent_t)
};

// Name of the ABI entry point currently executing (labels its elementwise launches in the profiler).
extern thread_local const char* tl_op_name;
struct OpName {
  explicit OpName(const char* n) : prev(tl_op_name) { if (!tl_op_name) tl_op_name = n; }
  ~OpName() { tl_op_name = prev; }
  const char* prev;
};

#define RB_LAUNCH_CHECK()                 \
  do {                                    \
    RB_CUDA_CHECK(cudaGetLastError());    \
    ::rrtmgpb::count_launch();            \
  } while (0)

void fused_set_constants(double grav, double m_dry);  // gas_optics_fused.cu
void table_cache_release(const void* key);            // gas_optics_fused.cu: drops the g-fastest copies keyed by a table
void table_cache_release_abi(const void* key);
// g-point-fastest copies of kmajor / kminor_* for the kernel-by-kernel ABI entry points; false: not available
// (cache switched off, see rrtmgpb_abi_table_cache)
// the extern symbols' tau_absorption / Planck source on the g-point-fastest kernels (gas_optics_fused.cu); device pointers
bool tau_absorption_gfast(int ncol, int nlay, int nbnd, int ngpt, int ngas, int nflav, int neta, int npres, int ntemp,
                          int nminorlower, int nminorklower, int nminorupper, int nminorkupper, int idx_h2o,
                          const int* gpoint_flavor, const int* band_lims_gpt, const Float* kmajor, const Float* kminor_lower,
                          const Float* kminor_upper, const int* minor_limits_gpt_lower, const int* minor_limits_gpt_upper,
                          const Bool* minor_scales_with_density_lower, const Bool* minor_scales_with_density_upper,
                          const Bool* scale_by_complement_lower, const Bool* scale_by_complement_upper,
                          const int* idx_minor_lower, const int* idx_minor_upper, const int* idx_minor_scaling_lower,
                          const int* idx_minor_scaling_upper, const int* kminor_start_lower, const int* kminor_start_upper,
                          const Bool* tropo, const Float* col_mix, const Float* fmajor, const Float* fminor, const Float* play,
                          const Float* tlay, const Float* col_gas, const int* jeta, const int* jtemp, const int* jpress,
                          Float* tau, bool accumulate);
bool planck_source_gfast(int ncol, int nlay, int nbnd, int ngpt, int nflav, int neta, int npres, int ntemp, int nPlanckTemp,
                         const Float* tlay, const Float* tlev, const Float* tsfc, int sfc_lay, const Float* fmajor,
                         const int* jeta, const Bool* tropo, const int* jtemp, const int* jpress, const int* band_lims_gpt,
                         const Float* pfracin, Float temp_ref_min, Float totplnk_delta, const Float* totplnk,
                         const int* gpoint_flavor, Float* sfc_src, Float* lay_src, Float* lev_src, Float* sfc_source_Jac);
bool tables_gfast_abi(const Float* kmajor, const Float* kminor_lower, const Float* kminor_upper, int ntemp, int neta,
                      int npres, int ngpt, int nkl, int nku, const Float** kmajorT, const Float** kminorT_lower,
                      const Float** kminorT_upper, int* gp, int* pitch_lower, int* pitch_upper);

// Express path (gas_optics_fused.cu -> solvers_abi.cu), per calling thread: the next broadband solver call ADDS its
// spectrally integrated fluxes to the output arrays (accumulate) and splits its g-points over `groups` grid rows, each
// row owning the copy of the outputs `group_stride` elements after the previous one.
struct ExpressSolverMode { int accumulate = 0; int groups = 1; size_t group_stride = 0; };
extern thread_local ExpressSolverMode tl_express;
// true while a library-internal driver (the express path) calls the ABI entry points with arrays it allocated itself:
// DevArg then skips the per-pointer provenance query (cudaPointerGetAttributes, ~1 us each, 20 per solver call -
// as much host time as the GPU needs for one of the express path's small launches)
extern thread_local bool tl_trust_device_ptrs;

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- argument staging ----------------------------------------------------------------------
enum class Dir { In, Out, InOut };

template <typename T>
class DevArg {
 public:
  // align: byte alignment the kernel's vector loads / stores assume for this array (0: element alignment).  A Fortran
  // host may pass array sections or arrays at their natural 4- / 8-byte alignment: a DEVICE pointer that does not meet
  // `align` is staged through an aligned scratch copy as well (device-to-device), instead of faulting in the kernel.
  DevArg(const T* p, size_t count, Dir dir, bool used = true, size_t align = 0)
      : host_(const_cast<T*>(p)), n_(count), dir_(dir), on_device_(false) {
    if (!used || p == nullptr || count == 0) { dev_ = const_cast<T*>(p); staged_ = false; return; }
    on_device_ = tl_trust_device_ptrs || is_device_ptr(p);
    if (on_device_ && (align == 0 || reinterpret_cast<uintptr_t>(p) % align == 0)) {
      dev_ = const_cast<T*>(p); staged_ = false; return;
    }
    staged_ = true;
    dev_ = static_cast<T*>(dev_alloc(n_ * sizeof(T)));
    if (dir_ != Dir::Out)
      RB_CUDA_CHECK(cudaMemcpyAsync(dev_, host_, n_ * sizeof(T), cudaMemcpyDefault, stream()));
  }
  ~DevArg() {
    if (!staged_) return;
    if (dir_ != Dir::In) {
      RB_CUDA_CHECK(cudaMemcpyAsync(host_, dev_, n_ * sizeof(T), cudaMemcpyDefault, stream()));
      if (!on_device_) RB_CUDA_CHECK(cudaStreamSynchronize(stream()));  // host results are valid at return
    }
    dev_free(dev_);
  }
  DevArg(const DevArg&) = delete;
  DevArg& operator=(const DevArg&) = delete;
  T* get() const { return dev_; }
  operator T*() const { return dev_; }
  bool staged() const { return staged_; }

 private:
  T* host_;
  T* dev_;
  size_t n_;
  Dir dir_;
  bool staged_;
  bool on_device_;
};

template <typename T> using In = DevArg<T>;

// numeric limits of the working precision (Fortran epsilon(), tiny())
#ifdef RTE_USE_SP
typedef float2 Float2;
#define make_Float2 make_float2
#else
typedef double2 Float2;
#define make_Float2 make_double2
#endif
#ifdef RTE_USE_SP
#define RB_EPS FLT_EPSILON
#define RB_TINY FLT_MIN
#else
#define RB_EPS DBL_EPSILON
#define RB_TINY DBL_MIN
#endif

}  // namespace rrtmgpb
