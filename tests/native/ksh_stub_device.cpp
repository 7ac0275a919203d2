// Device stubs for building csrc/host/ksh_host.cpp alone (no CUDA) under ASan/UBSan/TSan: every ks_* entry the host
// layer calls answers KS_ERR_NO_DEVICE, which is also what a packing-only context never reaches.  Test infrastructure.
#include "ksched.h"
#include <cstdio>
extern "C" {
void ks__set_error(const char*) {}
const char* ks_last_error(void) { return ""; }
int ks_snapshot_create(int, ks_snapshot**) { return KS_ERR_NO_DEVICE; }
void ks_snapshot_destroy(ks_snapshot*) {}
int ks_snapshot_set_nodes(ks_snapshot*, uint32_t, uint32_t, const int64_t*, const int64_t*, const uint64_t*) { return KS_ERR_NO_DEVICE; }
int ks_snapshot_set_bound(ks_snapshot*, uint64_t, const int32_t*, const int64_t*, const int64_t*) { return KS_ERR_NO_DEVICE; }
int ks_snapshot_apply_bind(ks_snapshot*, int32_t, int64_t, int64_t) { return KS_ERR_NO_DEVICE; }
int ks_check_cell(ks_snapshot*, int64_t, int64_t, const uint64_t*, uint32_t) { return KS_ERR_NO_DEVICE; }
int ks_select(ks_snapshot*, const ks_pods*, int, uint32_t, ks_bindings*, void*) { return KS_ERR_NO_DEVICE; }
int ks_stream_bind(ks_snapshot*, const ks_pods*, int, int32_t*, int64_t*, uint32_t*) { return KS_ERR_NO_DEVICE; }
int ks_select_sampling(ks_snapshot*, const ks_pods*, uint32_t, uint64_t, uint64_t, int32_t*, uint32_t*, int32_t*, uint8_t*) { return KS_ERR_NO_DEVICE; }
}
