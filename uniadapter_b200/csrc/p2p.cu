// All-gather of a small per-rank vector through NVLink peer memory, for the class-sharded cache step (every
// rank contributes 2 * K_p logits per sample). One CTA per rank does the whole exchange inside the step's stream / CUDA
// graph, so the step needs no collective library call (an NCCL all-gather is a host-side launch per step and could not
// be captured with the rest of the step in one graph):
//   1. push:   the rank's `n` floats are stored straight into slot [parity][rank] of EVERY peer's symmetric receive
//              buffer (peer pointers from torch.distributed._symmetric_memory: the buffers are mapped into this
//              process, the stores travel over NVLink);
//   2. signal: after a system-scope fence, flag [rank] of every peer is set to the step's sequence number
//              (st.release.sys);
//   3. wait:   thread r polls the local flag [r] until peer r's number has arrived (ld.acquire.sys, bounded: a peer that
//              never arrives raises the error word instead of hanging the GPU);
//   4. copy:   the gathered [P][n] block moves from the symmetric buffer into the step's ordinary buffer (volatile
//              loads: the lines were written by other GPUs).
// The receive buffer is double-buffered by the parity of the sequence number: a rank can run at most one exchange
// ahead of a peer (it needs the peer's flag of the current exchange to finish it), so two parities make the push of
// exchange i+1 safe while a slow peer still copies exchange i out. The sequence number lives on the device and is
// advanced by the kernel: a captured graph keeps counting.
#include "common.cuh"

namespace ua {
namespace {

__global__ void __launch_bounds__(256) p2p_allgather_kernel(const float* __restrict__ send, int n,
                                                            float* const* __restrict__ peer_recv,
                                                            int* const* __restrict__ peer_flag, int rank, int P,
                                                            int* __restrict__ seq_ptr, float* __restrict__ local_out,
                                                            int* __restrict__ err) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int seq = *seq_ptr + 1;
  const int parity = seq & 1;
  for (int r = 0; r < P; ++r) {
    float* dst = peer_recv[r] + ((size_t)parity * P + rank) * n;
    for (int i = tid; i < n; i += T) dst[i] = send[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < P) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flag[tid] + rank), "r"(seq) : "memory");
    const int* mine = peer_flag[rank] + tid;
    const long long t0 = clock64();
    int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > 6000000000LL) {      // ~3 s at 2 GHz: a missing peer must not hang the GPU
        atomicExch(err, 1 + tid);
        break;
      }
    }
  }
  __syncthreads();
  const float* src = peer_recv[rank] + (size_t)parity * P * n;
  for (int i = tid; i < P * n; i += T) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(src + i) : "memory");
    local_out[i] = v;
  }
  __syncthreads();
  if (tid == 0) *seq_ptr = seq;
}

}  // namespace
}  // namespace ua

extern "C" int ua_p2p_allgather_f32(const float* send, int n, const void* peer_recv_ptrs, const void* peer_flag_ptrs,
                                    int rank, int P, int* seq, float* local_out, int* err, void* stream) {
  using namespace ua;
  UA_REQUIRE(send && peer_recv_ptrs && peer_flag_ptrs && seq && local_out && err, "ua_p2p_allgather_f32: NULL pointer");
  UA_REQUIRE(n >= 1 && P >= 1 && P <= 256 && rank >= 0 && rank < P, "ua_p2p_allgather_f32: bad sizes n=%d P=%d rank=%d", n,
             P, rank);
  p2p_allgather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(send, n, (float* const*)peer_recv_ptrs,
                                                          (int* const*)peer_flag_ptrs, rank, P, seq, local_out, err);
  return check_launch("ua_p2p_allgather_f32");
}
