// builder.cu -- placeholder until the Voigt/line-binning kernels land (next milestone)
#include "builder.hpp"
namespace bart {
struct BuilderState { int dummy; };
void builder_run_and_write(BuilderState *&, const Options &, const Atmosphere &, const Molecules &,
                           Tli &, const std::vector<double> &, cudaStream_t, const std::string &p) {
  fail("opacity file '%s' does not exist and the grid builder is not available in this build", p.c_str());
}
void builder_slice(BuilderState *&, const Options &, const Atmosphere &, const Molecules &, Tli &,
                   const std::vector<double> &, cudaStream_t, int, int, double *) {
  fail("the grid builder is not available in this build");
}
long long builder_stats(BuilderState *, long long *, long long *, long long *) { return -1; }
long long builder_line_bins(BuilderState *, long long *, long long) { return -1; }
int builder_profile(BuilderState *, int, int, float *, long long, long long *) { return -1; }
void builder_free(BuilderState *b) { delete b; }
}  // namespace bart
