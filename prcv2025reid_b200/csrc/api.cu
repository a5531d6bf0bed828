// api.cu -- small C-ABI utilities of libreid_b200.
#include "common.cuh"

extern "C" size_t reid_pid_index_workspace_bytes(int64_t G);
extern "C" size_t reid_retrieve_fused_workspace_bytes(int64_t Q, int64_t G, int d);

extern "C" const char* reid_strerror(int code) {
  switch (code) {
    case REID_OK: return "ok";
    case REID_E_INVALID: return "invalid argument";
    case REID_E_CUDA: return "CUDA call failed";
    case REID_E_WORKSPACE: return "workspace too small";
    case REID_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

extern "C" int reid_abi_version(void) { return 2; }

extern "C" size_t reid_workspace_bytes(int which, int64_t Q, int64_t G, int d) {
  if (which == 0) return reid_pid_index_workspace_bytes(G);
  if (which == 1) return reid_retrieve_fused_workspace_bytes(Q, G, d);
  return 0;
}

extern "C" int reid_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}
