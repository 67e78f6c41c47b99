// Library-level entry points: version, error strings, per-device init.
#include "common.cuh"

namespace tgsr {
thread_local int g_last_cuda_error = 0;
int g_sm_count = 0;
int g_device = -1;
unsigned long long g_launches = 0;
int init_attention(void);   // attention.cu
int init_conv_tc(void);     // conv_tc.cu (tcgen05 path)
int init_blocked_misc(void);  // blocked_misc.cu
int init_lstm(void);          // lstm.cu
int init_attention_mma(void);  // attention_mma.cu
int init_attention_f32_mma(void);  // attention_f32_mma.cu
int init_attention_bwd(void);      // attention_bwd.cu
int init_damsm(void);              // damsm.cu
int init_conv_tf32(void);          // conv_tf32.cu
int init_pyramid(void);            // pyramid.cu
int init_wgrad_tc(void);           // wgrad_tc.cu
}  // namespace tgsr

extern "C" {

int tgsr_version(void) { return 100; }  // 0.1.0

const char* tgsr_error_string(int s) {
  switch (s) {
    case TGSR_OK: return "ok";
    case TGSR_ERR_BAD_SHAPE: return "bad shape or alignment";
    case TGSR_ERR_UNSUPPORTED: return "unsupported dimensions for this kernel";
    case TGSR_ERR_CUDA: return "CUDA error (see tgsr_last_cuda_error)";
    case TGSR_ERR_NULL: return "null pointer";
    case TGSR_ERR_NOT_INIT: return "tgsr_init(device) not called";
    default: return "unknown status";
  }
}

int tgsr_last_cuda_error(void) { return tgsr::g_last_cuda_error; }

int tgsr_init(int device) {
  TGSR_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  TGSR_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return TGSR_ERR_UNSUPPORTED;  // built for sm_100a only
  tgsr::g_sm_count = p.multiProcessorCount;
  tgsr::g_device = device;
  int s = tgsr::init_attention();
  if (s != TGSR_OK) return s;
  s = tgsr::init_conv_tc();
  if (s != TGSR_OK) return s;
  s = tgsr::init_blocked_misc();
  if (s != TGSR_OK) return s;
  s = tgsr::init_attention_mma();
  if (s != TGSR_OK) return s;
  s = tgsr::init_attention_f32_mma();
  if (s != TGSR_OK) return s;
  s = tgsr::init_attention_bwd();
  if (s != TGSR_OK) return s;
  s = tgsr::init_damsm();
  if (s != TGSR_OK) return s;
  s = tgsr::init_conv_tf32();
  if (s != TGSR_OK) return s;
  s = tgsr::init_pyramid();
  if (s != TGSR_OK) return s;
  s = tgsr::init_wgrad_tc();
  if (s != TGSR_OK) return s;
  return tgsr::init_lstm();
}

int tgsr_sm_count(void) { return tgsr::g_sm_count; }

unsigned long long tgsr_launch_count(void) { return tgsr::g_launches; }

}  // extern "C"
