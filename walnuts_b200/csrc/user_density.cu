// Model kind 5: the caller's own density as CUDA source, compiled at run time (NVRTC) INTO
// the chain-resident kernel.
//
// The reference takes any callable with the LogpGrad signature (concepts.hpp:25-60;
// LOGP_CFUNC over the C ABI, walnutpy.cpp:127-130) because its chains are host threads.
// A host function cannot feed a device batch, and the batched device callback (kind 4)
// runs on the lock-step engine, one HBM round trip of the state per gradient.  A density
// given as device source gets what the built-in targets get: it is instantiated as the
// `Target` of walnuts_chain_kernel (chain_kernel.cuh), so the chain stays in registers
// across the micro-steps of an orbit.  Two forms:
//
//   * element-wise (separable) densities -- logp(theta) = sum_d f_d(theta_d) -- define
//       __device__ void wb200_logp_grad(int d, double x, const double* par,
//                                       double& lp, double& g);
//     (term d of the log density and its derivative at x; par = the model's data1 doubles);
//   * anything that fits the Target interface (init / grad over a thread's K x 2 element
//     slots, with the group's all-reduce available, as FunnelTargetT does): define
//       template <int T, int K, class Real> struct MyTarget { ... };
//       #define WB200_USER_TARGET MyTarget
//
// The translation unit is  #include "engine_kernels.cuh"  + the source + a wrapper; the
// headers are the very files of this directory, embedded in the library at build time
// (build/embedded_headers.inc).  NVRTC and the driver API are loaded at run time, so the
// library has no link-time dependency on either and still loads where they are absent.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "engine.cuh"
#include "engine_shapes.cuh"
#include "user_density.cuh"

namespace {

#include "embedded_headers.inc"  // kEmbeddedNames[], kEmbeddedData[], kEmbeddedCount

struct Rtc {
  void* lib = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                               const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
};

struct Driver {
  void* lib = nullptr;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

template <class F>
void sym(void* lib, F& f, const char* name) {
  f = reinterpret_cast<F>(dlsym(lib, name));
  if (!f) throw std::runtime_error(std::string("symbol missing: ") + name);
}

Rtc& rtc() {
  static Rtc r;
  static std::once_flag once;
  static std::string failure;
  std::call_once(once, [] {
    const char* env = std::getenv("WB200_NVRTC");
    const char* names[] = {env, "libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) {
      if (n && !r.lib) r.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!r.lib) { failure = "libnvrtc.so.12 could not be loaded (set WB200_NVRTC)"; return; }
    try {
      sym(r.lib, r.CreateProgram, "nvrtcCreateProgram");
      sym(r.lib, r.DestroyProgram, "nvrtcDestroyProgram");
      sym(r.lib, r.AddNameExpression, "nvrtcAddNameExpression");
      sym(r.lib, r.CompileProgram, "nvrtcCompileProgram");
      sym(r.lib, r.GetProgramLogSize, "nvrtcGetProgramLogSize");
      sym(r.lib, r.GetProgramLog, "nvrtcGetProgramLog");
      sym(r.lib, r.GetLoweredName, "nvrtcGetLoweredName");
      sym(r.lib, r.GetCUBINSize, "nvrtcGetCUBINSize");
      sym(r.lib, r.GetCUBIN, "nvrtcGetCUBIN");
      sym(r.lib, r.GetErrorString, "nvrtcGetErrorString");
    } catch (const std::exception& e) {
      failure = e.what();
      r.lib = nullptr;
    }
  });
  if (!r.lib) throw std::runtime_error("run-time compilation unavailable: " + failure);
  return r;
}

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  static std::string failure;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!d.lib) { failure = "libcuda.so.1 could not be loaded"; return; }
    try {
      sym(d.lib, d.ModuleLoadData, "cuModuleLoadData");
      sym(d.lib, d.ModuleUnload, "cuModuleUnload");
      sym(d.lib, d.ModuleGetFunction, "cuModuleGetFunction");
      sym(d.lib, d.FuncSetAttribute, "cuFuncSetAttribute");
      sym(d.lib, d.OccupancyMaxActiveBlocksPerMultiprocessor,
          "cuOccupancyMaxActiveBlocksPerMultiprocessor");
      sym(d.lib, d.LaunchKernel, "cuLaunchKernel");
      sym(d.lib, d.GetErrorString, "cuGetErrorString");
    } catch (const std::exception& e) {
      failure = e.what();
      d.lib = nullptr;
    }
  });
  if (!d.lib) throw std::runtime_error("CUDA driver API unavailable: " + failure);
  return d;
}

void cu_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = nullptr;
  driver().GetErrorString(r, &msg);
  throw wb200::CudaError(std::string("CUDA driver error in ") + what + ": " +
                         (msg ? msg : "unknown"));
}

// what follows the caller's source: element-wise densities are wrapped into a Target
const char* kWrapper = R"WB200(
namespace wb200 {
#ifndef WB200_USER_TARGET
template <int T, int K, class Real>
struct UserElementwiseT {
  const double* par;
  int base, D;
  __device__ __forceinline__ void init(const ChainParams& p, int tid) {
    par = p.tparam; base = 2 * tid; D = p.D;
  }
  __device__ __forceinline__ void grad(const Real (&th)[K][2], Real (&g)[K][2],
                                       Real& lp_part, Group<T>&) const {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int d = base + 2 * k * T + v;
        double lp = 0.0, gg = 0.0;  // padding slots: no term, zero gradient
        if (d < D) ::wb200_logp_grad(d, static_cast<double>(th[k][v]), par, lp, gg);
        s += lp;
        g[k][v] = static_cast<Real>(gg);
      }
    }
    lp_part = static_cast<Real>(s);
  }
};
#define WB200_USER_TARGET UserElementwiseT
#endif
template <int T, int K> using UserTarget = WB200_USER_TARGET<T, K, double>;
}  // namespace wb200
)WB200";

struct Shape4 { int T, K, CTA, minb_adapt, minb_sample; };

// the launch shape a dimension gets (engine_shapes.cuh), with its register caps
Shape4 shape4_for(const wb200::LaunchShape& sh) {
  Shape4 out{};
#define WB200_TAKE_SHAPE(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_) \
  out = Shape4{T_, K_, CTA_, MINB_A_, MINB_S_}
  WB200_FOR_SHAPE(sh, WB200_TAKE_SHAPE, unused);
#undef WB200_TAKE_SHAPE
  return out;
}

// (heap-allocated and never destroyed: the modules live as long as the process, and no
// destructor calls into the driver while the process is being torn down)
std::mutex g_cache_mutex;
auto& g_cache = *new std::map<std::string, std::shared_ptr<wb200::UserModule>>();
// compiled once per (launch shape, source), loaded once per device
struct Compiled {
  std::vector<char> cubin;
  std::vector<std::string> lowered;
  std::string log;
};
auto& g_compiled = *new std::map<std::string, std::shared_ptr<Compiled>>();

}  // namespace

namespace wb200 {

UserModule::~UserModule() {
  // the context may already be gone at process exit; nothing to report then
  if (mod) driver().ModuleUnload(static_cast<CUmodule>(mod));
}

std::vector<char> user_compile(const char* source, const LaunchShape& shape,
                               std::vector<std::string>* lowered, std::string* log) {
  if (!source || !*source) throw std::invalid_argument("device source is empty");
  Rtc& r = rtc();
  const Shape4 s4 = shape4_for(shape);
  std::string src = "#include \"engine_kernels.cuh\"\n#line 1 \"device_source\"\n";
  src += source;
  src += "\n";
  src += kWrapper;
  nvrtcProgram prog = nullptr;
  nvrtcResult rc = r.CreateProgram(&prog, src.c_str(), "wb200_user_density.cu", kEmbeddedCount,
                                   kEmbeddedData, kEmbeddedNames);
  if (rc != NVRTC_SUCCESS) {
    throw std::runtime_error(std::string("nvrtcCreateProgram: ") + r.GetErrorString(rc));
  }
  struct Guard { Rtc& r; nvrtcProgram* p; ~Guard() { r.DestroyProgram(p); } } guard{r, &prog};
  auto tkc = [&](int minb, const char* adapt, const char* free_run) {
    std::stringstream ss;
    ss << "wb200::walnuts_chain_kernel<wb200::UserTarget<" << s4.T << "," << s4.K << ">,"
       << s4.T << "," << s4.K << "," << s4.CTA << "," << minb << "," << adapt << ",double,"
       << free_run << ">";
    return ss.str();
  };
  std::stringstream ik, ok;
  ik << "wb200::init_kernel<wb200::UserTarget," << s4.T << "," << s4.K << "," << s4.CTA << ">";
  ok << "wb200::orbit_kernel<wb200::UserTarget," << s4.T << "," << s4.K << "," << s4.CTA
     << ",double>";
  const std::vector<std::string> names = {
      tkc(s4.minb_adapt, "true", "false"), tkc(s4.minb_sample, "false", "false"),
      tkc(s4.minb_adapt, "true", "true"), tkc(s4.minb_sample, "false", "true"),
      ik.str(), ok.str()};
  for (const auto& n : names) {
    rc = r.AddNameExpression(prog, n.c_str());
    if (rc != NVRTC_SUCCESS) {
      throw std::runtime_error(std::string("nvrtcAddNameExpression: ") + r.GetErrorString(rc));
    }
  }
  // the flags of csrc/Makefile: the kernels' arithmetic is written without contraction
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo",
                        "-default-device"};
  rc = r.CompileProgram(prog, 5, opts);
  size_t log_size = 0;
  r.GetProgramLogSize(prog, &log_size);
  std::string text(log_size, '\0');
  if (log_size > 1) r.GetProgramLog(prog, text.data());
  while (!text.empty() && (text.back() == '\0' || text.back() == '\n')) text.pop_back();
  if (log) *log = text;
  if (rc != NVRTC_SUCCESS) {
    throw std::invalid_argument("the device source does not compile (" +
                                std::string(r.GetErrorString(rc)) + "):\n" + text);
  }
  if (lowered) {
    lowered->clear();
    for (const auto& n : names) {
      const char* low = nullptr;
      rc = r.GetLoweredName(prog, n.c_str(), &low);
      if (rc != NVRTC_SUCCESS || !low) {
        throw std::runtime_error("nvrtcGetLoweredName failed for " + n);
      }
      lowered->push_back(low);
    }
  }
  size_t size = 0;
  r.GetCUBINSize(prog, &size);
  std::vector<char> cubin(size);
  rc = r.GetCUBIN(prog, cubin.data());
  if (rc != NVRTC_SUCCESS || size == 0) {
    throw std::runtime_error(std::string("nvrtcGetCUBIN: ") + r.GetErrorString(rc));
  }
  return cubin;
}

std::shared_ptr<UserModule> user_module(const char* source, const LaunchShape& shape,
                                        int device) {
  std::stringstream ckey, key;
  ckey << shape.T << "x" << shape.K << ":" << source;
  key << device << ":" << ckey.str();
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_cache.find(key.str());
  if (it != g_cache.end()) return it->second;
  std::shared_ptr<Compiled>& comp = g_compiled[ckey.str()];
  if (!comp) {
    auto c = std::make_shared<Compiled>();
    c->cubin = user_compile(source, shape, &c->lowered, &c->log);
    comp = c;
  }
  const std::vector<std::string>& low = comp->lowered;
  const std::vector<char>& cubin = comp->cubin;
  auto m = std::make_shared<UserModule>();
  m->log = comp->log;
  Driver& d = driver();
  WB200_CUDA(cudaSetDevice(device));
  WB200_CUDA(cudaFree(nullptr));  // the primary context exists and is current
  CUmodule mod = nullptr;
  cu_check(d.ModuleLoadData(&mod, cubin.data()), "cuModuleLoadData");
  m->mod = mod;
  void** slots[] = {&m->adapt, &m->sample, &m->adapt_free, &m->sample_free, &m->init,
                    &m->orbit};
  for (int i = 0; i < 6; ++i) {
    CUfunction f = nullptr;
    cu_check(d.ModuleGetFunction(&f, mod, low[i].c_str()), "cuModuleGetFunction");
    *slots[i] = f;
  }
  g_cache[key.str()] = m;
  return m;
}

int user_blocks_per_sm(void* func, int cta, size_t dyn_smem) {
  Driver& d = driver();
  CUfunction f = static_cast<CUfunction>(func);
  if (dyn_smem > 48 * 1024) {
    cu_check(d.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                static_cast<int>(dyn_smem)), "cuFuncSetAttribute");
  }
  int n = 0;
  cu_check(d.OccupancyMaxActiveBlocksPerMultiprocessor(&n, f, cta, dyn_smem),
           "cuOccupancyMaxActiveBlocksPerMultiprocessor");
  return n > 1 ? n : 1;
}

void user_launch(void* func, int grid, int cta, size_t dyn_smem, cudaStream_t stream,
                 void* params) {
  void* args[] = {params};
  cu_check(driver().LaunchKernel(static_cast<CUfunction>(func), grid, 1, 1, cta, 1, 1,
                                 static_cast<unsigned>(dyn_smem),
                                 reinterpret_cast<CUstream>(stream), args, nullptr),
           "cuLaunchKernel");
}

}  // namespace wb200

extern "C" int wb200_compile_device_source(const char* source, int num_params, char* log,
                                           size_t log_size, WalnutpyError** err) {
  return wb200::catch_exceptions(err, [&] {
    if (num_params < 1) throw std::invalid_argument("num_params must be at least 1");
    std::string text;
    wb200::user_compile(source, wb200::shape_for_dim(num_params), nullptr, &text);
    if (log && log_size > 0) {
      const size_t n = std::min(log_size - 1, text.size());
      std::memcpy(log, text.data(), n);
      log[n] = '\0';
    }
  });
}
