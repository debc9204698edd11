// Image<->map association (placeholder until the tensor-core path lands in this file).
#include "ctx.cuh"
using namespace ppg;
namespace ppg {
struct AssocState {};
void assoc_destroy(ppg_ctx* c) { delete c->assoc; c->assoc = nullptr; }
}
extern "C" {
int ppg_upload_map(ppg_ctx* c, const float*, int) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_associate(ppg_ctx* c, const ppg_assoc_in*, ppg_assoc_out*) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_stage(ppg_ctx* c, const ppg_assoc_in*) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_run(ppg_ctx* c) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_fetch(ppg_ctx* c, ppg_assoc_out*) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_run_frame(ppg_ctx* c, int) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_fallback_rows(ppg_ctx* c, int*) { return set_err(c, PPG_ERR_ARG, "association not built"); }
int ppg_assoc_device_results(ppg_ctx* c, void**, void**, void**, void**, void**) { return set_err(c, PPG_ERR_ARG, "association not built"); }
}
