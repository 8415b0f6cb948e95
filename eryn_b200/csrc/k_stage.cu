// Part of eryn_b200 (kernel overview in common.cuh).
//
// K10: staging of a stored sample (Backend.save_step, backends/backend.py:1014-1091, call site ensemble.py:1013-1028).
//
// The reference copies the whole State into the backend arrays on the host at every stored step, NaN-filling the
// coordinates of inactive leaves (backend.py:1053-1059).  Here the walkers stay on the device between yields, so a
// stored step is: ONE pack kernel on the sampler's stream that gathers every array of the step (coords with the NaN mask
// applied, logl, logp, betas, leaf flags, accept mask, per-move accept counters, the control block with the swap counts)
// into one contiguous device staging slot — a snapshot, so the sampler can run on — followed by ONE device-to-host copy
// of the slot into pinned memory on a side stream.  The sampler's stream never waits for a store.
#include "common.cuh"

namespace eb {

struct StageArgs {
  int nseg;
  const unsigned char* src[EB_STAGE_MAX_SEGMENTS];
  unsigned long long off[EB_STAGE_MAX_SEGMENTS + 1];   // destination byte offsets (multiples of 8); off[nseg] = total
  unsigned long long nbytes[EB_STAGE_MAX_SEGMENTS];
  const uint8_t* mask_inds;    // segment 0 = coords: leaf flags [rows][L] (NULL: every leaf is stored as is)
  int mask_L, mask_D;
  double fill;                 // value stored for the coordinates of an inactive leaf
  unsigned char* dst;
};

__global__ void __launch_bounds__(256) stage_pack_kernel(const __grid_constant__ StageArgs p) {
  const unsigned long long total_words = p.off[p.nseg] >> 3;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total_words; w += stride) {
    const unsigned long long b = w << 3;
    int sidx = 0;
#pragma unroll 1
    while (sidx + 1 < p.nseg && b >= p.off[sidx + 1]) ++sidx;
    const unsigned long long rel = b - p.off[sidx];
    if (rel >= p.nbytes[sidx]) continue;                 // padding between segments
    const unsigned char* s = p.src[sidx] + rel;
    unsigned long long v;
    if (rel + 8 <= p.nbytes[sidx]) {
      v = *reinterpret_cast<const unsigned long long*>(s);
      if (sidx == 0 && p.mask_inds) {                    // backend.py:1053-1059: inactive leaves are stored as NaN
        const unsigned long long leaf = (rel >> 3) / (unsigned long long)p.mask_D;
        if (!p.mask_inds[leaf]) v = (unsigned long long)__double_as_longlong(p.fill);
      }
    } else {                                             // tail of a byte segment
      v = 0ull;
      for (unsigned long long i = 0; rel + i < p.nbytes[sidx]; ++i) v |= (unsigned long long)s[i] << (8 * i);
    }
    *reinterpret_cast<unsigned long long*>(p.dst + b) = v;
  }
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_stage_pack(const eb_stage* sg, void* stream) {
  if (!sg || !sg->dst) return fail(EB_ERR_INVALID, "stage description / destination is NULL");
  if (sg->nseg < 1 || sg->nseg > EB_STAGE_MAX_SEGMENTS) return fail(EB_ERR_INVALID, "nseg %d out of range", sg->nseg);
  StageArgs a;
  memset(&a, 0, sizeof(a));
  a.nseg = sg->nseg;
  unsigned long long o = 0;
  for (int i = 0; i < sg->nseg; ++i) {
    if (!sg->src[i] || sg->nbytes[i] == 0) return fail(EB_ERR_INVALID, "segment %d is empty", i);
    if ((reinterpret_cast<size_t>(sg->src[i]) & 7) != 0) return fail(EB_ERR_INVALID, "segment %d is not 8-byte aligned", i);
    a.src[i] = (const unsigned char*)sg->src[i];
    a.nbytes[i] = sg->nbytes[i];
    a.off[i] = o;
    o += (sg->nbytes[i] + 7ull) & ~7ull;
  }
  a.off[sg->nseg] = o;
  if (sg->dst_bytes < o) return fail(EB_ERR_INVALID, "staging slot of %llu bytes is too small (%llu needed)",
                                     (unsigned long long)sg->dst_bytes, o);
  a.mask_inds = sg->mask_inds;
  a.mask_L = sg->mask_nleaves; a.mask_D = sg->mask_ndim;
  if (a.mask_inds && (a.mask_L < 1 || a.mask_D < 1)) return fail(EB_ERR_INVALID, "mask needs nleaves/ndim");
  a.fill = sg->fill;
  a.dst = (unsigned char*)sg->dst;
  const unsigned long long words = o >> 3;
  unsigned long long g = (words + 255) / 256;
  if (g > 148ull * 8) g = 148ull * 8;     // grid-stride: a few CTAs per SM saturate HBM for a copy
  stage_pack_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("stage_pack");
}

}  // extern "C"
