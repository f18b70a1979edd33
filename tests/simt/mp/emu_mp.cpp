// TEST INFRASTRUCTURE ONLY: the mapper worker code of lra_b200/csrc/mp_*.cuh compiled for the CPU through the SIMT emulator.
// Built twice by tests/emu_mp.py: -DMP_LANES=32 (the product's lane count, lock-step fibers) and -DMP_LANES=1 (every warp
// collective is the identity; same source at CPU speed for bulk parity runs).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "mp_sdp_driver.cuh"

using namespace lra;
using namespace lra::mp;

extern "C" int emu_mp_lanes() { return MP_LANES; }

extern "C" int emu_sdp_batch(int n_prob, int max_aln, const int *mode, const uint64_t *frag_off, const uint32_t *q, const uint32_t *t, const int32_t *len,
                             const uint64_t *cl_off_off, const int *cl_off, const uint8_t *cl_strand, const int *only_cl, const float *rate, const int *irate,
                             const int *read_len, float alnthres, int NumAln, const int64_t *stops, const float *slope, const float *inter, int ceil1, int ceil2,
                             int *n_chains, int *chain_len, float *chain_val, uint32_t *bounds, uint32_t *chain, uint8_t *link, int *cl_of_frag,
                             uint64_t arena_bytes, uint64_t *peak) {
  Pwl pwl; for (int i = 0; i < 25; i++) { pwl.stops[i] = stops[i]; pwl.slope[i] = slope[i]; pwl.inter[i] = inter[i]; } pwl.ceil1 = ceil1; pwl.ceil2 = ceil2;
  std::vector<unsigned char> arena(arena_bytes + 64);
  int err = 0;
  SdpBatch b;
  b.n_prob = n_prob; b.max_aln = max_aln; b.mode = mode; b.frag_off = (const unsigned long long *)frag_off; b.q = q; b.t = t; b.len = len;
  b.cl_off_off = (const unsigned long long *)cl_off_off; b.cl_off = cl_off; b.cl_strand = cl_strand; b.only_cl = only_cl; b.rate = rate; b.irate = irate;
  b.read_len = read_len; b.alnthres = alnthres; b.NumAln = NumAln; b.pwl = &pwl;
  b.n_chains = n_chains; b.chain_len = chain_len; b.chain_val = chain_val; b.bounds = bounds; b.chain = chain; b.link = link; b.cl_of_frag = cl_of_frag;
  unsigned char *base = arena.data(); while (((uintptr_t)base) & 15) base++;
  b.arena = base; b.arena_per_warp = arena_bytes; b.err = &err; b.peak = (unsigned long long *)peak;
  emu::launch(dim3(1), dim3(MP_LANES), 0, [&] { sdp_batch_kernel(b); });
  return err;
}
