// pipeline.cu — the batched p7_Pipeline() driver (include/p7gpu.h §7).
#include "p7gpu_internal.h"

struct p7gpu_pipeline_result {
  p7gpu_pipeline_counters counters;
  std::vector<p7gpu_survivor> survivors;
  std::vector<float> fwd_xmx, bck_xmx;
};

extern "C" {

int p7gpu_pipeline_run(p7gpu_ctx *, const p7gpu_profile *, const p7gpu_seqdb *, const p7gpu_pipeline_opts *, p7gpu_pipeline_result **)
{
  p7g::set_error("p7gpu_pipeline_run: not implemented yet");
  return P7GPU_EINVAL;
}
const p7gpu_pipeline_counters *p7gpu_result_counters(const p7gpu_pipeline_result *res) { return res ? &res->counters : nullptr; }
int64_t p7gpu_result_nsurvivors(const p7gpu_pipeline_result *res) { return res ? (int64_t) res->survivors.size() : 0; }
const p7gpu_survivor *p7gpu_result_survivors(const p7gpu_pipeline_result *res) { return res ? res->survivors.data() : nullptr; }
const float *p7gpu_result_fwd_xmx(const p7gpu_pipeline_result *res) { return res ? res->fwd_xmx.data() : nullptr; }
const float *p7gpu_result_bck_xmx(const p7gpu_pipeline_result *res) { return res ? res->bck_xmx.data() : nullptr; }
void p7gpu_result_destroy(p7gpu_pipeline_result *res) { delete res; }

}
