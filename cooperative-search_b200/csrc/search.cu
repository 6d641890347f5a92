// placeholder until the search_env kernels land
#include "cs_common.cuh"
extern "C" {
int cs_search_create(const cs_search_cfg*, cs_search**) { cs_set_error("search_env not built yet"); return CS_ERR_UNSUPPORTED; }
void cs_search_destroy(cs_search*) {}
int cs_search_buffers_get(cs_search*, cs_search_buffers*) { return CS_ERR_UNSUPPORTED; }
int cs_search_env_info(const cs_search*, int32_t*) { return CS_ERR_UNSUPPORTED; }
int cs_search_set_targets(cs_search*, const int32_t*, void*) { return CS_ERR_UNSUPPORTED; }
int cs_search_reset(cs_search*, const uint8_t*, uint32_t, void*) { return CS_ERR_UNSUPPORTED; }
int cs_search_step(cs_search*, const uint8_t*, void*) { return CS_ERR_UNSUPPORTED; }
int cs_search_step_random(cs_search*, int32_t, void*) { return CS_ERR_UNSUPPORTED; }
int cs_search_step_host(cs_search*, const cs_search_host_io*, void*) { return CS_ERR_UNSUPPORTED; }
int cs_search_stats(cs_search*, double*, void*) { return CS_ERR_UNSUPPORTED; }
}
