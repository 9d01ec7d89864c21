// Library identity and error strings.
#include "common.cuh"

extern "C" int as_abi_version(void) { return 1; }

extern "C" int as_compiled_sm(void) {
#ifdef AS_COMPILED_SM
  return AS_COMPILED_SM;
#else
  return 0;
#endif
}

extern "C" const char* as_error_string(int code) {
  switch (code) {
    case AS_OK: return "ok";
    case AS_ERR_BAD_ARG: return "anystereo_b200: bad argument (null pointer or non-positive size)";
    case AS_ERR_UNSUPPORTED: return "anystereo_b200: unsupported shape/option";
    case AS_ERR_INDEX_RANGE: return "anystereo_b200: tensor exceeds 32-bit indexing";
    case AS_ERR_ALIGNMENT: return "anystereo_b200: pointer/pitch alignment violated";
    case AS_ERR_DRIVER: return "anystereo_b200: CUDA driver entry point unavailable";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "anystereo_b200: unknown error";
}

static int g_operand_fmt = AS_FMT_BF16;
int as_operand_f16_internal() { return g_operand_fmt != AS_FMT_BF16; }
int as_operand_fmt_internal() { return g_operand_fmt; }

extern "C" int as_set_operand_format(int fmt) {
  if (fmt != AS_FMT_BF16 && fmt != AS_FMT_F16 && fmt != AS_FMT_F16F8) return AS_ERR_BAD_ARG;
  g_operand_fmt = fmt;
  return AS_OK;
}
extern "C" int as_get_operand_format(void) { return g_operand_fmt; }
