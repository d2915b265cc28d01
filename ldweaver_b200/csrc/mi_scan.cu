#include "../../include/ldw.h"
#include "ctx.h"
using namespace ldw;
extern "C" {
int ldw_mi_plan_create(ldw_ctx*, const uint8_t*, int64_t, int64_t, const double*, const int32_t*, const int32_t*, int64_t, ldw_mi_plan**) { return set_error(LDW_ERR_UNSUPPORTED, "stub"); }
void ldw_mi_plan_destroy(ldw_mi_plan*) {}
int ldw_mi_scan(ldw_mi_plan*, double, double, double, double, int, int, int, ldw_links*, ldw_links*, ldw_links*, double*, double*, ldw_scan_stats*) { return set_error(LDW_ERR_UNSUPPORTED, "stub"); }
int ldw_mi_block_dense(ldw_mi_plan*, int64_t, double*, int64_t*, int64_t*) { return set_error(LDW_ERR_UNSUPPORTED, "stub"); }
int ldw_mi_pairs_exact(ldw_mi_plan*, int64_t, const int32_t*, const int32_t*, int64_t, double*) { return set_error(LDW_ERR_UNSUPPORTED, "stub"); }
}
