/* pidnet_b200 -- hardware probes (tools/probe_*.py).  NOT part of the product library: these entry points exist only in
 * libpidnet_b200_probe.so (`python -m pidnet_b200.build --probes`), which is the product sources + csrc/probe.cu compiled with
 * -DPIDNET_PROBES.  They document how tcgen05 reads shifted windows of a TMA-written halo patch, MN-major operands and the
 * CTA-pair (cta_group::2) operand split, and measure MMA issue rates. */
#ifndef PIDNET_B200_PROBE_H_
#define PIDNET_B200_PROBE_H_
#ifdef __cplusplus
extern "C" {
#endif

/* Hardware probe used by tools/probe_halo.py (documents how tcgen05 reads shifted windows of a
 * TMA-written halo patch; not on the product path). */
int pidnet_probe_halo(void* stream, const void* x_18x10x64_bf16, const void* w_64x64_bf16, int r, int s, int mode,
                      float* out_128x64);

int pidnet_probe_mn(void* stream, const void* a_64x128_bf16, const void* b_64x64_bf16, int lbo, int sbo, float* out_128x64);
int pidnet_probe_mma_rate(void* stream, int N, int iters, int distinct, int blocks, long long* out_cycles_dev);
/* CTA-pair (tcgen05 cta_group::2, M = 256) probes: operand-split convention and MMA rate (tools/probe_pair.py). */
int pidnet_probe_pair(void* stream, const void* a_256x64_bf16, const void* b_64x64_bf16, int swap_b, float* out_256x64);
int pidnet_probe_mma_rate_pair(void* stream, int N, int iters, int distinct, int pairs, long long* out_cycles_dev);

#ifdef __cplusplus
}
#endif
#endif /* PIDNET_B200_PROBE_H_ */
