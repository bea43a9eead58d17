/* kvb200_test.h — TEST-ONLY additions of libkvb200_test.so (the product library libkvb200.so exports none of these).
 *
 * libkvb200_test.so is libkvb200.so's own objects plus two groups of hooks compiled with -DKVB_TEST_HOOKS:
 *   - a mock CUDA VMM driver, so that the VMM engine's host logic is testable without a GPU
 *     (tests/test_vmm_engine_cpu.py);
 *   - per-call driver cost probes (tools/probe_vmm*.py; cf. the reference's benchmarks/bench_vmm/bench_vmm.cpp).
 * It also exports everything include/kvb200.h declares.
 */
#ifndef KVB200_TEST_H_
#define KVB200_TEST_H_

#include "kvb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Host-logic test aid: replaces the CUDA VMM driver calls with legality-checking stand-ins (map only onto
 * unmapped reserved space, unmap / set-access only whole mappings, release only unmapped handles) that count calls
 * and model a device of `total_bytes`, so the engine's decisions are testable without a GPU.  Must be called before
 * the first VMM call of the process.  out10 = {creates, releases, maps, unmaps, set-access, violations, live
 * handles, live mappings, (unused), handle bytes}. */
int kvb_vmm_mock_driver(int64_t total_bytes, int latency_us);
int kvb_vmm_mock_set_other_bytes(int64_t bytes);   /* memory taken by "somebody else" */
int kvb_vmm_mock_stats(int64_t *out10);
const char *kvb_vmm_mock_last_violation(void);

/* per-call driver costs in microseconds (measurement aid, cf. benchmarks/bench_vmm/bench_vmm.cpp):
 * out8 = {cuMemCreate, cuMemMap, cuMemSetAccess per unit, cuMemSetAccess whole range (one call),
 *         cuMemUnmap per unit, cuMemUnmap whole range (one call), cuMemRelease, unmap+map+setaccess per unit} */
int kvb_vmm_probe(int device_index, int64_t unit_bytes, int64_t n_units, int64_t iters, double *out8);
/* map/unmap cycle cost against M live zero-page mappings: mode bit0 = one cuMemSetAccess over the whole range
 * at init instead of one per unit, bit1 = cycle units of a separate small reservation.
 * out3 = {us per map cycle, us per unmap cycle, init ms} */
int kvb_vmm_probe2(int device_index, int64_t unit_bytes, int64_t m_units, int mode, int64_t iters, double *out3);
/* cycle cost against M live physical mappings: inside the same reservation / in a separate one / in one placed
 * directly behind it.  out7 = {same_map, same_unmap, separate_map, separate_unmap, adjacent_map, adjacent_unmap, adjacent?} */
int kvb_vmm_probe3(int device_index, int64_t unit_bytes, int64_t m_units, int64_t iters, double *out7);

#ifdef __cplusplus
}
#endif
#endif /* KVB200_TEST_H_ */
