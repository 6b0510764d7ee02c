// placeholder until the tcgen05 tower lands (next commit)
#include "../../include/caro_b200.h"
#include "common_host.h"
#include "net.h"
int caro_net_tc_pack(caro_net*, const float*) { return CARO_OK; }
void caro_net_tc_free(caro_net*) {}
int caro_net_tc_forward(caro_net*, int, int, int, const void*, const uint8_t*, const int32_t*, int64_t, float*, float*, cudaStream_t) {
  return caro_fail(CARO_E_STATE, "tcgen05 tower not built");
}
