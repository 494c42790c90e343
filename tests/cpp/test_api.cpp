// TEST: the C++ entry point walnuts_b200::walnuts (walnuts_b200/host/api.hpp) driven the
// way examples/walnutpie_api.cpp:49-68 drives walnutpie::walnuts.
//   test_api cpu            handler-count and config errors, no-GPU failure (runs anywhere)
//   test_api gpu <out.bin>  a run on the GPU; writes [C][warm+samp][D] draws for pytest
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../walnuts_b200/host/api.hpp"

namespace wb = walnuts_b200;

struct ChainRecorder {  // concepts.hpp:212-245
  std::vector<double> warm, samp, lps;
  std::vector<double> final_inv_mass;
  double final_step = 0;
  int warmup_complete_calls = 0;
  void on_warmup(const wb::Vector& position, double lp, double step_size,
                 const wb::Vector& diag_inv_mass) {
    warm.insert(warm.end(), position.begin(), position.end());
    lps.push_back(lp);
    if (!(step_size > 0) || diag_inv_mass.size() != position.size()) std::abort();
  }
  void on_warmup_complete(double step_size, const wb::Vector& diag_inv_mass) {
    final_step = step_size;
    final_inv_mass = diag_inv_mass;
    ++warmup_complete_calls;
  }
  void on_sample(const wb::Vector& position, double lp) {
    samp.insert(samp.end(), position.begin(), position.end());
    lps.push_back(lp);
  }
  void on_logp_exception(const wb::Vector&, const std::exception&) noexcept {}
};
struct RhatRecorder {  // concepts.hpp:174-176
  std::vector<double> values;
  void on_r_hat(double r) { values.push_back(r); }
};
struct NeverInterrupt {  // concepts.hpp:186-188
  void throw_if_interrupted() const {}
};
struct InterruptAfter {
  mutable int calls = 0;
  int limit;
  void throw_if_interrupted() const {
    if (++calls > limit) throw std::runtime_error("interrupted");
  }
};

static wb::WalnutsConfig make_config(std::size_t C, std::size_t D, std::size_t warm,
                                     std::size_t samp) {
  auto init = wb::InitConfigBuilder(C, D).step_sizes(0.4).build();
  auto warmup = wb::WarmupConfigBuilder().min_max_iter(warm, warm).build();
  auto sampling = wb::SamplingConfigBuilder().min_max_iter(samp, samp).build();
  return wb::WalnutsConfig(std::move(init), std::move(warmup), std::move(sampling));
}

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::cerr << "FAILED: " #cond << " at line " << __LINE__ << std::endl; \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  const std::size_t C = 6, D = 5, warm = 40, samp = 30;
  WalnutModelDesc model{0, static_cast<int>(D), 0, nullptr, nullptr};  // std normal
  RhatRecorder rhat;
  NeverInterrupt never;
  if (mode == "cpu") {
    std::vector<ChainRecorder> few(C - 1);
    try {  // api.hpp:41-44
      wb::walnuts(1, few, rhat, never, model, make_config(C, D, warm, samp));
      EXPECT(false);
    } catch (const std::invalid_argument& e) {
      EXPECT(std::string(e.what()) ==
             "chain_handlers.size() must be equal to config.init().num_chains()");
    }
    try {  // config.hpp:650-656
      wb::WarmupConfigBuilder().min_max_iter(5, 2);
      EXPECT(false);
    } catch (const std::invalid_argument& e) {
      EXPECT(std::string(e.what()).find("min_iter cannot be greater") != std::string::npos);
    }
    std::vector<ChainRecorder> handlers(C);
    try {
      wb::walnuts(1, handlers, rhat, never, model, make_config(C, D, warm, samp));
      std::cout << "ran on a GPU" << std::endl;
    } catch (const std::runtime_error& e) {  // no CPU path
      EXPECT(std::string(e.what()).find("no CUDA device") != std::string::npos);
      std::cout << "no GPU: " << e.what() << std::endl;
    }
    std::cout << "cpu checks passed" << std::endl;
    return 0;
  }
  // ---- gpu
  std::vector<ChainRecorder> handlers(C);
  auto cfg = make_config(C, D, warm, samp);
  wb::DeviceInit dev;
  dev.random_positions = true;
  dev.init_scale = 1.5;
  wb::walnuts(1234, handlers, rhat, never, model, cfg, dev);
  double mean = 0, m2 = 0;
  for (auto& h : handlers) {
    EXPECT(h.warm.size() == warm * D);
    EXPECT(h.samp.size() == samp * D);
    EXPECT(h.lps.size() == warm + samp);
    EXPECT(h.warmup_complete_calls == 1);
    EXPECT(h.final_step > 0 && h.final_inv_mass.size() == D);
    for (double x : h.samp) { mean += x; m2 += x * x; }
  }
  const double n = static_cast<double>(C * samp * D);
  mean /= n;
  EXPECT(std::abs(mean) < 0.3 && std::abs(m2 / n - 1.0) < 0.4);  // N(0, 1) marginals
  // adaptive lengths: the R-hat controller reports and may stop early
  {
    std::vector<ChainRecorder> h2(C);
    RhatRecorder r2;
    auto init = wb::InitConfigBuilder(C, D).step_sizes(0.4).build();
    auto warmup = wb::WarmupConfigBuilder().min_max_iter(20, 60).build();
    auto sampling = wb::SamplingConfigBuilder().min_max_iter(10, 400).build();
    wb::walnuts(7, h2, r2, never, model,
                wb::WalnutsConfig(std::move(init), std::move(warmup), std::move(sampling)),
                dev);
    EXPECT(!r2.values.empty());
    const std::size_t got = h2[0].samp.size() / D;
    EXPECT(got >= 10 && got <= 400);
    for (auto& h : h2) EXPECT(h.samp.size() == got * D);  // all chains stop together
    EXPECT(got == 400 || r2.values.back() <= 1.01);
  }
  // the interrupt callback ends the run (adapt.hpp:227)
  {
    std::vector<ChainRecorder> h3(C);
    InterruptAfter stop{0, 3};
    try {
      wb::walnuts(7, h3, rhat, stop, model, make_config(C, D, warm, samp), dev);
      EXPECT(false);
    } catch (const std::runtime_error& e) {
      EXPECT(std::string(e.what()) == "interrupted");
    }
    EXPECT(h3[0].warm.size() == 4 * 5 * D);  // four blocks of publish_stride = 5 published
  }
  if (argc > 2) {
    std::ofstream out(argv[2], std::ios::binary);
    for (auto& h : handlers) {
      out.write(reinterpret_cast<const char*>(h.warm.data()), h.warm.size() * 8);
      out.write(reinterpret_cast<const char*>(h.samp.data()), h.samp.size() * 8);
    }
  }
  std::cout << "gpu checks passed" << std::endl;
  return 0;
}
