// b3d — neighbour halo exchange and small all-reduces of depth-slab sharded inference (slab.py; BASELINE config 4)
// as plain kernels over NVLink PEER MEMORY: every rank owns one symmetric buffer (same layout on every GPU, mapped
// into all peers), senders store boundary slices straight into the neighbour's mailbox and publish a sequence
// number with a system-scope release store, receivers spin on their own flag with acquire loads.  No NCCL call, no
// host synchronisation, CUDA-graph capturable; ~3 us per kernel instead of ~50 us per NCCL send/recv group.
//
// Symmetric buffer layout (bytes; identical on every rank):
//   [    0,  2048)  u64 flags:  [dir*2 + slot] halo (dir 0 = written by rank-1, dir 1 = written by rank+1);
//                                [8 + slot*16 + r] all-reduce contribution of rank r
//   [ 2048,  4096)  u32 local "blocks done" counters (one per direction), never touched by peers
//   [ 4096, 69632)  all-reduce slots  [slot 2][rank 16][256 x 8 bytes]
//   [131072, ... )  mailboxes         [dir 2][slot 2][mailbox_bytes]
// Sequence numbers: value = epoch * 65536 + seq + 1 where `epoch` is a device counter ticked once per forward (so a
// replayed CUDA graph publishes fresh values) and `seq` counts the exchanges inside one forward.  Slots alternate
// with seq; every exchange is bidirectional at the protocol level (a flag is sent even when no data moves), hence a
// rank that has completed exchange s knows both neighbours have consumed exchange s-1 and slot s%2 may be reused at
// s+1 ... s+2 is safe by induction.
#include "common.cuh"

namespace b3d {

constexpr int kOffDone = 2048, kOffAr = 4096, kOffMailbox = 131072;
constexpr int kArMaxRanks = 16, kArMaxWords = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until *flag >= want; a lost message becomes a trap after ~4 s instead of a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}

struct HaloSide {
  const float4* src;          // boundary slice to send (nullptr / n16 = 0: flag only)
  char* peer;                 // base of the neighbour's symmetric buffer (nullptr: no neighbour on this side)
  float4* dst;                // where the received slice goes (recv kernel)
  long long n16_send, n16_recv;
  int peer_dir;               // direction index the NEIGHBOUR files my message under
};

// grid (blocks, 2): y = side (0: towards rank-1, 1: towards rank+1)
__global__ void __launch_bounds__(256)
    halo_send_kernel(HaloSide s0, HaloSide s1, char* self, const unsigned long long* epoch, int seq,
                     long long mailbox_bytes) {
  const HaloSide s = blockIdx.y == 0 ? s0 : s1;
  if (s.peer == nullptr) return;
  const int slot = seq & 1;
  float4* mb = reinterpret_cast<float4*>(s.peer + kOffMailbox + (long long)(s.peer_dir * 2 + slot) * mailbox_bytes);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < s.n16_send; i += (long long)gridDim.x * 256)
    mb[i] = s.src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(self + kOffDone) + blockIdx.y;
    if (atomicAdd(done, 1u) == gridDim.x - 1) {       // last block of this side: publish
      *done = 0u;
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(s.peer) + s.peer_dir * 2 + slot,
                     *epoch * 65536ULL + (unsigned long long)seq + 1ULL);
    }
  }
}

// grid (blocks, 2): y = side the message comes FROM (0: rank-1, 1: rank+1)
__global__ void __launch_bounds__(256)
    halo_recv_kernel(HaloSide s0, HaloSide s1, char* self, const unsigned long long* epoch, int seq,
                     long long mailbox_bytes) {
  const HaloSide s = blockIdx.y == 0 ? s0 : s1;
  if (s.peer == nullptr) return;
  const int slot = seq & 1, dir = blockIdx.y;
  if (threadIdx.x == 0)
    wait_flag(reinterpret_cast<const unsigned long long*>(self) + dir * 2 + slot,
              *epoch * 65536ULL + (unsigned long long)seq + 1ULL);
  __syncthreads();
  const float4* mb = reinterpret_cast<const float4*>(self + kOffMailbox + (long long)(dir * 2 + slot) * mailbox_bytes);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < s.n16_recv; i += (long long)gridDim.x * 256) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mb + i));
    s.dst[i] = v;
  }
}

// send + receive in ONE launch: every CTA first stores its share of the outgoing slice into the neighbour's mailbox
// (the last one to finish publishes the flag), then waits for the neighbour's flag and copies its share of the incoming
// slice.  Sends never wait for anything, and the <= 128 CTAs of the grid are co-resident, so the spin cannot deadlock.
__global__ void __launch_bounds__(256)
    halo_xchg_kernel(HaloSide s0, HaloSide s1, char* self, const unsigned long long* epoch, int seq,
                     long long mailbox_bytes) {
  const HaloSide s = blockIdx.y == 0 ? s0 : s1;
  if (s.peer == nullptr) return;
  const int slot = seq & 1, dir = blockIdx.y;
  const unsigned long long want = *epoch * 65536ULL + (unsigned long long)seq + 1ULL;
  float4* mbo = reinterpret_cast<float4*>(s.peer + kOffMailbox + (long long)(s.peer_dir * 2 + slot) * mailbox_bytes);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < s.n16_send; i += (long long)gridDim.x * 256)
    mbo[i] = s.src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* done = reinterpret_cast<unsigned int*>(self + kOffDone) + blockIdx.y;
    if (atomicAdd(done, 1u) == gridDim.x - 1) {       // last block of this side: publish
      *done = 0u;
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(s.peer) + s.peer_dir * 2 + slot, want);
    }
    wait_flag(reinterpret_cast<const unsigned long long*>(self) + dir * 2 + slot, want);
  }
  __syncthreads();
  const float4* mbi = reinterpret_cast<const float4*>(self + kOffMailbox + (long long)(dir * 2 + slot) * mailbox_bytes);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < s.n16_recv; i += (long long)gridDim.x * 256) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mbi + i));
    s.dst[i] = v;
  }
}

// one CTA: x (n <= 64 words of T) is summed over all ranks, in rank order (bit-identical on every rank)
template <typename T>
__global__ void __launch_bounds__(256)
    peer_allreduce_kernel(T* x, int n, const long long* peers, int rank, int world, char* self,
                          const unsigned long long* epoch, int seq) {
  const int slot = seq & 1;
  const unsigned long long want = *epoch * 65536ULL + (unsigned long long)seq + 1ULL;
  // scatter my contribution into slot[rank] of every rank (incl. myself)
  for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
    const int p = i / n, e = i - p * n;
    T* dst = reinterpret_cast<T*>(reinterpret_cast<char*>(peers[p]) + kOffAr +
                                  (long long)((slot * kArMaxRanks + rank) * kArMaxWords) * 8);
    dst[e] = x[e];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys(reinterpret_cast<unsigned long long*>(peers[threadIdx.x]) + 8 + slot * kArMaxRanks + rank, want);
  if (threadIdx.x < world)
    wait_flag(reinterpret_cast<const unsigned long long*>(self) + 8 + slot * kArMaxRanks + threadIdx.x, want);
  __syncthreads();
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    T acc = T(0);
    for (int r = 0; r < world; ++r) {
      const volatile T* src = reinterpret_cast<const volatile T*>(self + kOffAr +
                                                                  (long long)((slot * kArMaxRanks + r) * kArMaxWords) * 8);
      acc += src[e];
    }
    x[e] = acc;
  }
}

// the same for a PAIR of vectors in one exchange: `a` (fp64, na values: GroupNorm chunk statistics) and `b` (fp32, nb
// values: SE pooling sums) — a ResnetBlock's conv2 statistics and its pooling sums are needed at the same point
__global__ void __launch_bounds__(256)
    peer_allreduce2_kernel(double* a, int na, float* b, int nb, const long long* peers, int rank, int world, char* self,
                           const unsigned long long* epoch, int seq) {
  const int slot = seq & 1;
  const unsigned long long want = *epoch * 65536ULL + (unsigned long long)seq + 1ULL;
  const long long my_off = kOffAr + (long long)((slot * kArMaxRanks + rank) * kArMaxWords) * 8;
  for (int i = threadIdx.x; i < (na + nb) * world; i += blockDim.x) {
    const int p = i / (na + nb), e = i - p * (na + nb);
    char* dst = reinterpret_cast<char*>(peers[p]) + my_off;
    if (e < na) reinterpret_cast<double*>(dst)[e] = a[e];
    else reinterpret_cast<float*>(dst + (long long)na * 8)[e - na] = b[e - na];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)
    st_release_sys(reinterpret_cast<unsigned long long*>(peers[threadIdx.x]) + 8 + slot * kArMaxRanks + rank, want);
  if (threadIdx.x < world)
    wait_flag(reinterpret_cast<const unsigned long long*>(self) + 8 + slot * kArMaxRanks + threadIdx.x, want);
  __syncthreads();
  for (int e = threadIdx.x; e < na + nb; e += blockDim.x) {
    if (e < na) {
      double acc = 0.0;
      for (int r = 0; r < world; ++r)
        acc += reinterpret_cast<const volatile double*>(self + kOffAr + (long long)((slot * kArMaxRanks + r) * kArMaxWords) * 8)[e];
      a[e] = acc;
    } else {
      float acc = 0.f;
      for (int r = 0; r < world; ++r)
        acc += reinterpret_cast<const volatile float*>(self + kOffAr + (long long)((slot * kArMaxRanks + r) * kArMaxWords) * 8 +
                                                       (long long)na * 8)[e - na];
      b[e - na] = acc;
    }
  }
}

__global__ void epoch_tick_kernel(unsigned long long* epoch) { *epoch += 1ULL; }

static int bytes_view(const DLTensor* t, const char* name, TView* v, long long* nbytes) {
  B3D_TRY(view(t, DT_F32, -1, false, name, v));
  *nbytes = v->numel * 4;
  B3D_REQUIRE((*nbytes % 16) == 0 && (((uintptr_t)v->p) & 15) == 0, B3D_ERR_LAYOUT,
              "%s: halo slices must be 16-byte aligned multiples of 16 bytes", name);
  return B3D_OK;
}

}  // namespace b3d

using namespace b3d;

extern "C" long long b3d_slab_sym_bytes(long long mailbox_bytes) { return kOffMailbox + 4 * mailbox_bytes; }

// One neighbour exchange of a depth slab (slab.py: SlabContext.with_halo).  send_prev / send_next: my first / last
// boundary slices (fp32, contiguous; nullable = nothing to move that way), recv_prev / recv_next: where the
// neighbours' slices go (nullable).  prev_base / next_base: device addresses of the neighbours' symmetric buffers
// (0 = no neighbour).  sym: my symmetric buffer (uint8 viewed as fp32 here), epoch: int64 [1] device counter.
extern "C" int b3d_halo_exchange(const DLTensor* send_prev_, const DLTensor* send_next_, DLTensor* recv_prev_,
                                 DLTensor* recv_next_, long long prev_base, long long next_base, DLTensor* sym_,
                                 const DLTensor* epoch_, int seq, long long mailbox_bytes, void* stream) {
  TView sym, ep, t;
  B3D_TRY(view(sym_, DT_F32, 1, false, "sym", &sym));
  B3D_TRY(view(epoch_, DT_I64, 1, false, "epoch", &ep));
  B3D_REQUIRE(sym.numel * 4 >= b3d_slab_sym_bytes(mailbox_bytes) && mailbox_bytes % 16 == 0, B3D_ERR_SHAPE,
              "halo_exchange: symmetric buffer too small for the mailbox size");
  HaloSide s[2];
  memset(s, 0, sizeof(s));
  const DLTensor* snd[2] = {send_prev_, send_next_};
  DLTensor* rcv[2] = {recv_prev_, recv_next_};
  const long long base[2] = {prev_base, next_base};
  long long max16 = 1;
  for (int i = 0; i < 2; ++i) {
    s[i].peer = reinterpret_cast<char*>(base[i]);
    s[i].peer_dir = 1 - i;     // my message towards rank-1 is, for that rank, the one "from rank+1" (dir 1)
    long long nb = 0;
    if (snd[i] != nullptr && base[i] != 0) {
      B3D_TRY(bytes_view(snd[i], "halo send", &t, &nb));
      B3D_REQUIRE(nb <= mailbox_bytes, B3D_ERR_SHAPE, "halo_exchange: slice (%lld B) exceeds the mailbox (%lld B)", nb,
                  mailbox_bytes);
      s[i].src = reinterpret_cast<const float4*>(t.p);
      s[i].n16_send = nb / 16;
    }
    if (rcv[i] != nullptr && base[i] != 0) {
      B3D_TRY(bytes_view(rcv[i], "halo recv", &t, &nb));
      B3D_REQUIRE(nb <= mailbox_bytes, B3D_ERR_SHAPE, "halo_exchange: slice exceeds the mailbox");
      s[i].dst = reinterpret_cast<float4*>(t.p);
      s[i].n16_recv = nb / 16;
    }
    if (s[i].n16_send > max16) max16 = s[i].n16_send;
    if (s[i].n16_recv > max16) max16 = s[i].n16_recv;
  }
  if (base[0] == 0 && base[1] == 0) return B3D_OK;
  long long blocks = (max16 + 256 * 8 - 1) / (256 * 8);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)blocks, 2, 1);
  halo_xchg_kernel<<<grid, 256, 0, st>>>(s[0], s[1], (char*)sym.p, (const unsigned long long*)ep.p, seq, mailbox_bytes);
  B3D_LAUNCH_CHECK("halo_xchg");
  return B3D_OK;
}

// In-place sum over all ranks of a small vector (fp32 or fp64, <= 64 eight-byte words: GroupNorm chunk statistics,
// SE pooling sums).  peers: int64 [world] device addresses of every rank's symmetric buffer.
extern "C" int b3d_peer_allreduce(DLTensor* x_, const DLTensor* peers_, int rank, DLTensor* sym_,
                                  const DLTensor* epoch_, int seq, void* stream) {
  TView pe, sym, ep, x;
  B3D_TRY(view(peers_, DT_I64, 1, false, "peers", &pe));
  B3D_TRY(view(sym_, DT_F32, 1, false, "sym", &sym));
  B3D_TRY(view(epoch_, DT_I64, 1, false, "epoch", &ep));
  const int world = (int)pe.numel;
  B3D_REQUIRE(world >= 1 && world <= kArMaxRanks && rank >= 0 && rank < world, B3D_ERR_ARG, "peer_allreduce: bad world/rank");
  cudaStream_t st = (cudaStream_t)stream;
  const bool f64 = x_->dtype.code == kDLFloat && x_->dtype.bits == 64;
  B3D_TRY(view(x_, f64 ? DT_F64 : DT_F32, -1, false, "x", &x));
  B3D_REQUIRE(x.numel * (f64 ? 8 : 4) <= kArMaxWords * 8, B3D_ERR_SHAPE, "peer_allreduce: at most %d bytes", kArMaxWords * 8);
  if (f64)
    peer_allreduce_kernel<double><<<1, 256, 0, st>>>((double*)x.p, (int)x.numel, (const long long*)pe.p, rank, world,
                                                     (char*)sym.p, (const unsigned long long*)ep.p, seq);
  else
    peer_allreduce_kernel<float><<<1, 256, 0, st>>>((float*)x.p, (int)x.numel, (const long long*)pe.p, rank, world,
                                                    (char*)sym.p, (const unsigned long long*)ep.p, seq);
  B3D_LAUNCH_CHECK("peer_allreduce");
  return B3D_OK;
}

// In-place sums over all ranks of a fp64 vector and a fp32 vector in ONE exchange (together <= 2 KB).
extern "C" int b3d_peer_allreduce2(DLTensor* a_, DLTensor* b_, const DLTensor* peers_, int rank, DLTensor* sym_,
                                   const DLTensor* epoch_, int seq, void* stream) {
  TView pe, sym, ep, a, b;
  B3D_TRY(view(peers_, DT_I64, 1, false, "peers", &pe));
  B3D_TRY(view(sym_, DT_F32, 1, false, "sym", &sym));
  B3D_TRY(view(epoch_, DT_I64, 1, false, "epoch", &ep));
  B3D_TRY(view(a_, DT_F64, -1, false, "a", &a));
  B3D_TRY(view(b_, DT_F32, -1, false, "b", &b));
  const int world = (int)pe.numel;
  B3D_REQUIRE(world >= 1 && world <= kArMaxRanks && rank >= 0 && rank < world, B3D_ERR_ARG, "peer_allreduce: bad world/rank");
  B3D_REQUIRE(a.numel * 8 + b.numel * 4 <= kArMaxWords * 8, B3D_ERR_SHAPE, "peer_allreduce2: at most %d bytes", kArMaxWords * 8);
  peer_allreduce2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((double*)a.p, (int)a.numel, (float*)b.p, (int)b.numel,
                                                             (const long long*)pe.p, rank, world, (char*)sym.p,
                                                             (const unsigned long long*)ep.p, seq);
  B3D_LAUNCH_CHECK("peer_allreduce2");
  return B3D_OK;
}

extern "C" int b3d_epoch_tick(DLTensor* epoch_, void* stream) {
  TView ep;
  B3D_TRY(view(epoch_, DT_I64, 1, false, "epoch", &ep));
  epoch_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)ep.p);
  B3D_LAUNCH_CHECK("epoch_tick");
  return B3D_OK;
}
