// forge_ec_cuda.cu -- host side of the C ABI: context, device set-up, sharding over devices, the
// host-buffer pipeline and the extern "C" entry points declared in include/forge_ec_cuda.h.
// The kernels live in k_*.cu (see kernels.cuh).  There is no CPU code path in this library: every
// entry point needs a CUDA device of compute capability 10.x.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "kdf.cuh"
#include "kernels.cuh"

// ======================================================================================
// host side: context, sharding, C ABI
// ======================================================================================
using namespace fec;

namespace {

// host-buffer pipeline granule (a multiple of every tile size).  FORGE_EC_CHUNK_LG=16..22 overrides it (measurements).
static size_t chunk_lanes() {
  static const size_t v = [] {
    const char* e = getenv("FORGE_EC_CHUNK_LG");
    const int lg = e != nullptr ? atoi(e) : 0;
    // capacity of a pipeline slot.  2^19 against 2^18: 58.3 -> 59.3 M/s end to end (fewer partial waves); r2: 2^20 for calls of
    // 2^21 lanes and more, 64.1 -> 64.6 M/s (profiles/r2al_e2e_chunk.log); smaller calls keep 2^19-lane chunks (chunk_step)
    return size_t(1) << ((lg >= 16 && lg <= 22) ? lg : 20);
  }();
  return v;
}
#define CHUNK_LANES (chunk_lanes())
// lanes per chunk of one call on one device: the slot capacity for large calls, at most 2^19 below 2^21 lanes (so that a call of
// 2^20 lanes still overlaps its copies with its kernels)
static size_t chunk_step(size_t count) {
  const size_t cap = CHUNK_LANES, small = size_t(1) << 19;
  return (count >= (size_t(1) << 21) || cap <= small) ? cap : small;
}
constexpr int N_PIPE = 3;                        // in-flight chunks per device

struct Device {
  int id = -1;
  int sms = 0;
  u32* comb[2] = {nullptr, nullptr};  // secp256k1, P-256
  u32* comb_ct[2] = {nullptr, nullptr};  // the 4-bit combs of the constant-time table form (32 KiB each)
  u32* niels_ct = nullptr;               // the same for Ed25519 (48 KiB)
  u32* niels = nullptr;
  cudaStream_t streams[N_PIPE] = {};
  uint8_t* buf_in_a[N_PIPE] = {};   // scalars / seeds / field a
  uint8_t* buf_in_b[N_PIPE] = {};   // points / u / field b
  uint8_t* buf_out[N_PIPE] = {};
  uint8_t* buf_st[N_PIPE] = {};
  uint8_t* buf_x[N_PIPE] = {};      // ECDH: the 32-byte shared secrets before key derivation
  bool pipe_ready = false;
  // pinned staging for callers that pass PAGEABLE host memory (a Rust Vec<u8>, a numpy array): in a, in b, out, x, status per slot
  uint8_t* stage[N_PIPE] = {};
  bool stage_ready = false;
};
constexpr size_t STAGE_OFF_B = 64, STAGE_OFF_OUT = 128, STAGE_OFF_X = 128 + 66, STAGE_OFF_ST = 128 + 66 + 32, STAGE_BYTES_PER_LANE = 128 + 66 + 32 + 1;

}  // namespace

struct fec_ctx {
  // per-stage timing of the secp256k1 k*P kernels (fec_ctx_stage_timing): events of the calls made while enabled
  bool stage_timing = false;
  std::vector<cudaEvent_t> stage_ev;
  int stage_used = 0;
  std::vector<Device> devs;
  std::mutex mu;       // serialises the host-buffer entry points (they own the per-device pipeline buffers)
  std::mutex err_mu;   // the per-device host threads of one call may fail at the same time
  std::atomic<uint64_t> launches{0};
  char last_err[256] = {0};
};

namespace {

void set_err(fec_ctx* ctx, const char* file, int line, const char* what, const char* msg) {
  std::lock_guard<std::mutex> lock(ctx->err_mu);
  snprintf(ctx->last_err, sizeof(ctx->last_err), "%s:%d %s: %s", file, line, what, msg);
}
#define FEC_CUDA(ctx, call)                                                                     \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      set_err((ctx), __FILE__, __LINE__, #call, cudaGetErrorString(e_));                        \
      return e_ == cudaErrorMemoryAllocation ? FEC_ERR_OOM : FEC_ERR_CUDA;                      \
    }                                                                                           \
  } while (0)

// scope guards for the early-return paths of FEC_CUDA: the caller's device is restored, stream-ordered scratch is given back,
// and a device's pipeline streams are drained before the caller may touch (or free) its buffers again
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct AsyncBuf {
  uint8_t* p = nullptr;
  cudaStream_t st = nullptr;
  ~AsyncBuf() { if (p != nullptr) cudaFreeAsync(p, st); }
};
struct DrainGuard {
  Device& d;
  explicit DrainGuard(Device& dev) : d(dev) {}
  ~DrainGuard();
};

int device_init(fec_ctx* ctx, Device& d) {
  FEC_CUDA(ctx, cudaSetDevice(d.id));
  cudaDeviceProp prop;
  FEC_CUDA(ctx, cudaGetDeviceProperties(&prop, d.id));
  if (prop.major < 10) {
    char msg[96];
    snprintf(msg, sizeof(msg), "device %d is sm_%d%d; this library is built for sm_100a only", d.id, prop.major, prop.minor);
    set_err(ctx, __FILE__, __LINE__, "device_init", msg);
    return FEC_ERR_NO_DEVICE;
  }
  d.sms = prop.multiProcessorCount;
  FEC_CUDA(ctx, wmul_var_setup());
  FEC_CUDA(ctx, wmul_fixed_setup());
  FEC_CUDA(ctx, ed25519_setup());
  FEC_CUDA(ctx, x25519_vm_setup());
  for (int s = 0; s < N_PIPE; s++) FEC_CUDA(ctx, cudaStreamCreateWithFlags(&d.streams[s], cudaStreamNonBlocking));
  FEC_CUDA(ctx, cudaMalloc(&d.comb[0], COMB_WORDS * 4));
  FEC_CUDA(ctx, cudaMalloc(&d.comb[1], COMB_WORDS * 4));
  FEC_CUDA(ctx, cudaMalloc(&d.niels, NIELS_WORDS * 4));
  u32* bases = nullptr;  // window bases, only needed while the tables are built
  FEC_CUDA(ctx, cudaMalloc(&bases, (size_t)COMB_WINDOWS * 32 * 4 * 3));
  FEC_CUDA(ctx, build_comb_launch(FEC_SECP256K1, d.streams[0], d.comb[0], bases));
  FEC_CUDA(ctx, build_comb_launch(FEC_P256, d.streams[0], d.comb[1], bases + COMB_WINDOWS * 32));
  FEC_CUDA(ctx, build_niels_launch(d.streams[0], d.niels, bases + COMB_WINDOWS * 64));
  for (int c = 0; c < 2; c++) {
    FEC_CUDA(ctx, cudaMalloc(&d.comb_ct[c], (size_t)CT_TABLE_WORDS * 4));
    FEC_CUDA(ctx, build_comb_ct_launch(c, d.streams[0], d.comb_ct[c]));
  }
  FEC_CUDA(ctx, cudaMalloc(&d.niels_ct, (size_t)CT_NIELS_WORDS * 4));
  FEC_CUDA(ctx, build_niels_ct_launch(d.streams[0], d.niels_ct));
  ctx->launches += 9;
  FEC_CUDA(ctx, cudaStreamSynchronize(d.streams[0]));
  cudaFree(bases);
  return FEC_OK;
}

int device_pipe_init(fec_ctx* ctx, Device& d) {
  if (d.pipe_ready) return FEC_OK;
  FEC_CUDA(ctx, cudaSetDevice(d.id));
  for (int s = 0; s < N_PIPE; s++) {
    FEC_CUDA(ctx, cudaMalloc(&d.buf_in_a[s], CHUNK_LANES * 32));
    FEC_CUDA(ctx, cudaMalloc(&d.buf_in_b[s], CHUNK_LANES * 64));
    FEC_CUDA(ctx, cudaMalloc(&d.buf_out[s], CHUNK_LANES * 65 + 64));
    FEC_CUDA(ctx, cudaMalloc(&d.buf_st[s], CHUNK_LANES));
    FEC_CUDA(ctx, cudaMalloc(&d.buf_x[s], CHUNK_LANES * 32));
  }
  d.pipe_ready = true;
  return FEC_OK;
}

// cudaMemcpyAsync on pageable host memory is staged through the driver's own pinned buffer and BLOCKS the calling thread, so the
// three-stream pipeline of run_shard degenerates to copy -> compute -> copy (measured: 59.4 -> 35.8 M/s end to end on 2^24 lanes).
// For such buffers the library stages through its own pinned memory: the host thread copies chunk i+1 in (and chunk i-2 out)
// while the GPU works on the chunks in between.  FORGE_EC_STAGE=0 disables it (A/B).
int device_stage_init(fec_ctx* ctx, Device& d) {
  if (d.stage_ready) return FEC_OK;
  for (int s = 0; s < N_PIPE; s++) FEC_CUDA(ctx, cudaHostAlloc(&d.stage[s], CHUNK_LANES * STAGE_BYTES_PER_LANE + 256, cudaHostAllocDefault));
  d.stage_ready = true;
  return FEC_OK;
}
bool is_pageable(const void* p) {
  if (p == nullptr) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
bool staging_enabled() {
  static const bool v = [] { const char* e = getenv("FORGE_EC_STAGE"); return !(e != nullptr && e[0] == '0'); }();
  return v;
}

void device_free(Device& d) {
  if (d.id < 0) return;
  cudaSetDevice(d.id);
  for (int s = 0; s < N_PIPE; s++) {
    if (d.stage[s]) cudaFreeHost(d.stage[s]);
    if (d.streams[s]) cudaStreamSynchronize(d.streams[s]);
    cudaFree(d.buf_in_a[s]); cudaFree(d.buf_in_b[s]); cudaFree(d.buf_out[s]); cudaFree(d.buf_st[s]); cudaFree(d.buf_x[s]);
    if (d.streams[s]) cudaStreamDestroy(d.streams[s]);
  }
  cudaFree(d.comb[0]); cudaFree(d.comb[1]); cudaFree(d.niels); cudaFree(d.comb_ct[0]); cudaFree(d.comb_ct[1]); cudaFree(d.niels_ct);
  scratch_trim(d.id);  // the library's pool gives its cached scratch back to the driver
}


DrainGuard::~DrainGuard() {
  for (int s = 0; s < N_PIPE; s++) if (d.streams[s]) cudaStreamSynchronize(d.streams[s]);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- launches on device pointers ------------------------------------------------------
int launch_mul(fec_ctx* ctx, Device& d, int curve, int sem, const uint8_t* ds, const uint8_t* dp, int pfmt, size_t n, int ofmt,
               uint8_t* dout, uint8_t* dst, cudaStream_t st) {
  if (n == 0) return FEC_OK;
  if (sem == FEC_SEM_FORGE_QUIRK) {  // validated by the entry points: affine X||Y points only
    ctx->launches++;
    FEC_CUDA(ctx, curve == FEC_SECP256K1 ? quirk_k1_launch(st, ds, dp, n, ofmt, dout, dst) : quirk_p256_launch(st, ds, dp, n, ofmt, dout, dst));
  } else if (dp == nullptr) {
    ctx->launches++;
    if (fixed_ct_table()) FEC_CUDA(ctx, wmul_fixed_ct_launch(curve, d.sms, st, ds, d.comb_ct[curve], n, ofmt, dout, dst));
    else FEC_CUDA(ctx, wmul_fixed_launch(curve, d.sms, st, ds, d.comb[curve], n, ofmt, dout, dst));
  } else {
    int launches = 0;
    StageEvents se{nullptr, 0, 0}, *sep = nullptr;
    if (ctx->stage_timing && curve == FEC_SECP256K1) {
      const int need = 4 * (int)((n + GLV_CHUNK_LANES - 1) / GLV_CHUNK_LANES);
      while ((int)ctx->stage_ev.size() < ctx->stage_used + need) {
        cudaEvent_t ev;
        FEC_CUDA(ctx, cudaEventCreate(&ev));
        ctx->stage_ev.push_back(ev);
      }
      se = StageEvents{ctx->stage_ev.data() + ctx->stage_used, need, 0};
      sep = &se;
    }
    const cudaError_t e = wmul_var_launch(curve, d.sms, st, ds, dp, pfmt, n, ofmt, dout, dst, &launches, sep);
    if (sep != nullptr) ctx->stage_used += se.used / 4 * 4;
    ctx->launches += (uint64_t)launches;
    FEC_CUDA(ctx, e);
  }
  return FEC_OK;
}
// x(k*P) into `dx` (32 B per lane) with the lane status into `dst` (always needed: dead lanes give zero records), then the
// key derivation into `dout`.  kdf == KDF_NONE: the shared secrets are the output.
int launch_ecdh(fec_ctx* ctx, Device& d, int curve, const uint8_t* ds, const uint8_t* dp, int pfmt, size_t n, int kdf,
                const uint8_t* info, int info_len, int out_len, uint8_t* dout, uint8_t* dst, uint8_t* dx, cudaStream_t st) {
  if (n == 0) return FEC_OK;
  int rc = launch_mul(ctx, d, curve, FEC_SEM_STD, ds, dp, pfmt, n, 3, kdf == KDF_NONE ? dout : dx, dst, st);
  if (rc || kdf == KDF_NONE) return rc;
  ctx->launches++;
  FEC_CUDA(ctx, kdf_launch(st, kdf, dx, dst, n, info, info_len, out_len, dout));
  return FEC_OK;
}
static bool x_use_vm() {  // FORGE_EC_X_IMPL=vm: the accumulator-machine form of the ladder kernel (A/B: 45.0 against 48.9 M/s)
  static const bool v = [] { const char* e = getenv("FORGE_EC_X_IMPL"); return e != nullptr && e[0] == 'v'; }();
  return v;
}
int launch_x25519(fec_ctx* ctx, Device& d, int sem, const uint8_t* ds, const uint8_t* du, size_t n, uint8_t* dout, cudaStream_t st) {
  if (n == 0) return FEC_OK;
  ctx->launches++;
  if (sem == FEC_SEM_FORGE_QUIRK)  // the reference's literal x25519(), one lane per thread (k_quirk_k1.cu)
    FEC_CUDA(ctx, quirk_x25519_launch(st, ds, du, n, dout));
  else if (du == nullptr)  // base point u = 9: fixed-base path on the Edwards form of the curve (k_ed25519.cu, mode 2)
    FEC_CUDA(ctx, fixed_ct_table() ? ed25519_ct_launch(d.sms, st, ds, 2, d.niels_ct, n, dout) : ed25519_launch(d.sms, st, ds, 2, d.niels, n, dout));
  else if (x_use_vm())
    FEC_CUDA(ctx, x25519_vm_launch(d.sms, st, ds, du, n, dout));
  else
    FEC_CUDA(ctx, x25519_launch(d.sms, st, ds, du, n, dout));
  return FEC_OK;
}
int launch_ed(fec_ctx* ctx, Device& d, int sem, const uint8_t* din, int is_seed, size_t n, uint8_t* dout, cudaStream_t st) {
  if (n == 0) return FEC_OK;
  ctx->launches++;
  if (sem == FEC_SEM_FORGE_QUIRK)  // the reference's literal derive_public_key (seeds only), one lane per thread
    FEC_CUDA(ctx, quirk_ed25519_launch(st, din, n, dout));
  else
    FEC_CUDA(ctx, fixed_ct_table() ? ed25519_ct_launch(d.sms, st, din, is_seed ? 1 : 0, d.niels_ct, n, dout)
                                   : ed25519_launch(d.sms, st, din, is_seed ? 1 : 0, d.niels, n, dout));
  return FEC_OK;
}
// signing-side nonce commitment (k_nonce.cu).  Weierstrass modes: R = k G as affine X || Y into `dxy` with its lane status into
// `dst_tmp`, then the tail kernel (r = x(R) / the even-y rule) writes dR, dr and dst (dst may be dst_tmp).  EdDSA: the 64-byte
// value mod L into dr, then [r]B from the fixed-base table into dR.
int launch_nonce(fec_ctx* ctx, Device& d, int curve, int mode, const uint8_t* dnonce, size_t n, uint8_t* dR, uint8_t* dr, uint8_t* dst,
                 uint8_t* dxy, uint8_t* dst_tmp, cudaStream_t st) {
  if (n == 0) return FEC_OK;
  if (mode == FEC_NONCE_EDDSA) {
    ctx->launches++;
    FEC_CUDA(ctx, sc_reduce64_launch(st, dnonce, n, dr));
    if (dst != nullptr) FEC_CUDA(ctx, cudaMemsetAsync(dst, 0, n, st));
    return launch_ed(ctx, d, FEC_SEM_STD, dr, 0, n, dR, st);
  }
  int rc = launch_mul(ctx, d, curve, FEC_SEM_STD, dnonce, nullptr, FEC_PT_AFFINE_XY_BE, n, FEC_OUT_AFFINE_XY_BE, dxy, dst_tmp, st);
  if (rc) return rc;
  ctx->launches++;
  FEC_CUDA(ctx, nonce_tail_launch(curve, mode, st, dxy, dnonce, dst_tmp, n, dR, dr, dst));
  return FEC_OK;
}
int launch_field(fec_ctx* ctx, Device& d, int prime, int op, size_t lanes, int iters, const uint8_t* a, const uint8_t* b,
                 uint8_t* o, cudaStream_t st) {
  if (lanes == 0) return FEC_OK;
  ctx->launches++;
  FEC_CUDA(ctx, field_bench_launch(d.sms, st, prime, op, a, b, o, lanes, iters));
  return FEC_OK;
}

// ---- host-buffer pipeline: H2D -> kernel -> D2H in CHUNK_LANES granules over N_PIPE streams ----
struct HostJob {
  int kind;  // 0 mul, 1 x25519, 2 ed25519, 3 ecdh (mul + key derivation), 4 nonce commitment (two outputs)
  int curve, sem, pfmt, ofmt, is_seed;
  const uint8_t* in_a; size_t rec_a;  // kind 4: the nonces (32 or 64 B per lane)
  const uint8_t* in_b; size_t rec_b;
  uint8_t* out; size_t rec_out;
  uint8_t* status;
  int kdf = 0, info_len = 0, out_len = 0;  // kind 3 only
  const uint8_t* info = nullptr;
  int mode = 0;                            // kind 4 only
  uint8_t* out2 = nullptr;                 // kind 4: out_r, 32 B per lane
};
int launch_nonce(fec_ctx* ctx, Device& d, int curve, int mode, const uint8_t* dnonce, size_t n, uint8_t* dR, uint8_t* dr, uint8_t* dst,
                 uint8_t* dxy, uint8_t* dst_tmp, cudaStream_t st);

int run_shard(fec_ctx* ctx, Device& d, const HostJob& job, size_t first, size_t count) {
  if (count == 0) return FEC_OK;
  int rc = device_pipe_init(ctx, d);
  if (rc) return rc;
  FEC_CUDA(ctx, cudaSetDevice(d.id));
  DrainGuard drain(d);  // whatever happens below, no copy into the caller's buffers is in flight when this function returns
  // pageable caller memory goes through the device's pinned staging slots (see device_stage_init)
  const bool staged = staging_enabled() && count > 0 &&
                      (is_pageable(job.in_a) || is_pageable(job.in_b) || is_pageable(job.out) || is_pageable(job.out2) || is_pageable(job.status));
  if (staged && (rc = device_stage_init(ctx, d)) != FEC_OK) return rc;
  struct Pending { size_t g0 = 0, m = 0; bool live = false; } pend[N_PIPE];
  auto flush = [&](int s) {  // a finished chunk's results: staging -> the caller's buffers
    if (!pend[s].live) return;
    const size_t g0 = pend[s].g0, m = pend[s].m;
    memcpy(job.out + g0 * job.rec_out, d.stage[s] + CHUNK_LANES * STAGE_OFF_OUT, m * job.rec_out);
    if (job.out2) memcpy(job.out2 + g0 * 32, d.stage[s] + CHUNK_LANES * STAGE_OFF_X, m * 32);
    if (job.status) memcpy(job.status + g0, d.stage[s] + CHUNK_LANES * STAGE_OFF_ST, m);
    pend[s].live = false;
  };
  int slot = 0;
  const size_t step = chunk_step(count);
  for (size_t off = 0; off < count; off += step, slot = (slot + 1) % N_PIPE) {
    const size_t m = (count - off) < step ? (count - off) : step;
    const size_t g0 = first + off;
    cudaStream_t st = d.streams[slot];
    FEC_CUDA(ctx, cudaStreamSynchronize(st));  // slot's previous chunk fully drained
    const uint8_t* ha = job.in_a + g0 * job.rec_a;
    const uint8_t* hb = job.in_b != nullptr ? job.in_b + g0 * job.rec_b : nullptr;
    if (staged) {
      flush(slot);
      memcpy(d.stage[slot], ha, m * job.rec_a);
      ha = d.stage[slot];
      if (hb != nullptr) {
        memcpy(d.stage[slot] + CHUNK_LANES * STAGE_OFF_B, hb, m * job.rec_b);
        hb = d.stage[slot] + CHUNK_LANES * STAGE_OFF_B;
      }
    }
    uint8_t* da = job.rec_a > 32 ? d.buf_in_b[slot] : d.buf_in_a[slot];  // 64-byte input records live in the larger buffer
    FEC_CUDA(ctx, cudaMemcpyAsync(da, ha, m * job.rec_a, cudaMemcpyHostToDevice, st));
    const uint8_t* db = nullptr;
    if (hb != nullptr) {
      FEC_CUDA(ctx, cudaMemcpyAsync(d.buf_in_b[slot], hb, m * job.rec_b, cudaMemcpyHostToDevice, st));
      db = d.buf_in_b[slot];
    }
    if (job.kind == 0)
      rc = launch_mul(ctx, d, job.curve, job.sem, d.buf_in_a[slot], db, job.pfmt, m, job.ofmt, d.buf_out[slot],
                      job.status ? d.buf_st[slot] : nullptr, st);
    else if (job.kind == 3)
      rc = launch_ecdh(ctx, d, job.curve, d.buf_in_a[slot], db, job.pfmt, m, job.kdf, job.info, job.info_len, job.out_len,
                       d.buf_out[slot], d.buf_st[slot], d.buf_x[slot], st);
    else if (job.kind == 1)
      rc = launch_x25519(ctx, d, job.sem, d.buf_in_a[slot], db, m, d.buf_out[slot], st);
    else if (job.kind == 5) {  // point decoding: the records arrive in buf_in_b (33 or 64 B per lane)
      ctx->launches++;
      const cudaError_t e = point_decode_launch(job.curve, st, da, job.pfmt, m, job.ofmt, d.buf_out[slot], job.status ? d.buf_st[slot] : nullptr);
      if (e != cudaSuccess) { set_err(ctx, __FILE__, __LINE__, "point_decode_launch", cudaGetErrorString(e)); rc = FEC_ERR_CUDA; }
    } else if (job.kind == 4)  // R -> buf_out, r -> buf_x; the affine R of the Weierstrass modes passes through buf_in_b (unused there)
      rc = launch_nonce(ctx, d, job.curve, job.mode, da, m, d.buf_out[slot], d.buf_x[slot], d.buf_st[slot], d.buf_in_b[slot],
                        d.buf_st[slot], st);
    else
      rc = launch_ed(ctx, d, job.sem, d.buf_in_a[slot], job.is_seed, m, d.buf_out[slot], st);
    if (rc) return rc;
    uint8_t* ho = staged ? d.stage[slot] + CHUNK_LANES * STAGE_OFF_OUT : job.out + g0 * job.rec_out;
    uint8_t* hx = staged ? d.stage[slot] + CHUNK_LANES * STAGE_OFF_X : (job.out2 ? job.out2 + g0 * 32 : nullptr);
    uint8_t* hs = staged ? d.stage[slot] + CHUNK_LANES * STAGE_OFF_ST : (job.status ? job.status + g0 : nullptr);
    FEC_CUDA(ctx, cudaMemcpyAsync(ho, d.buf_out[slot], m * job.rec_out, cudaMemcpyDeviceToHost, st));
    if (job.out2) FEC_CUDA(ctx, cudaMemcpyAsync(hx, d.buf_x[slot], m * 32, cudaMemcpyDeviceToHost, st));
    if (job.status) FEC_CUDA(ctx, cudaMemcpyAsync(hs, d.buf_st[slot], m, cudaMemcpyDeviceToHost, st));
    if (staged) { pend[slot].g0 = g0; pend[slot].m = m; pend[slot].live = true; }
  }
  for (int k = 0; k < N_PIPE; k++) {  // oldest slot first
    const int s = (slot + k) % N_PIPE;
    FEC_CUDA(ctx, cudaStreamSynchronize(d.streams[s]));
    if (staged) flush(s);
  }
  return FEC_OK;
}

// contiguous lane ranges over the context's devices, one host thread per device
int run_host_job(fec_ctx* ctx, const HostJob& job, size_t n) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t G = ctx->devs.size();
  if (G == 1) return run_shard(ctx, ctx->devs[0], job, 0, n);
  std::vector<size_t> cut(G + 1, 0);
  for (size_t i = 0; i < G; i++) {
    size_t first = 0, count = 0;
    fec_shard_range(n, (int)G, (int)i, &first, &count);
    cut[i] = first;
    cut[i + 1] = first + count;
  }
  std::vector<int> rcs(G, FEC_OK);
  std::vector<std::thread> th;
  for (size_t i = 0; i < G; i++)
    th.emplace_back([&, i]() { rcs[i] = run_shard(ctx, ctx->devs[i], job, cut[i], cut[i + 1] - cut[i]); });
  for (auto& t : th) t.join();
  for (int rc : rcs) if (rc) return rc;
  return FEC_OK;
}

// FEC_SEM_FORGE_QUIRK replays the reference's literal limb code; restated for both Weierstrass curves, affine X||Y input only
// (the reference's own SEC1 decoder rejects its own encodings, test_summary.md:77-85, so there is nothing to replay)
bool sem_supported(int sem, int curve, int points_fmt, int out_fmt) {
  if (sem == FEC_SEM_STD) return true;
  // (the literal replay encodes compressed / uncompressed / X||Y records only)
  return sem == FEC_SEM_FORGE_QUIRK && (curve == FEC_SECP256K1 || curve == FEC_P256) && points_fmt == FEC_PT_AFFINE_XY_BE &&
         out_fmt <= FEC_OUT_AFFINE_XY_BE;
}

Device* dev_at(fec_ctx* ctx, int dev_index) {
  if (ctx == nullptr || dev_index < 0 || (size_t)dev_index >= ctx->devs.size()) return nullptr;
  return &ctx->devs[(size_t)dev_index];
}

}  // namespace

extern "C" {

const char* fec_strerror(int err) {
  switch (err) {
    case FEC_OK: return "ok";
    case FEC_ERR_INVALID_ARG: return "invalid argument";
    case FEC_ERR_CUDA: return "CUDA runtime error";
    case FEC_ERR_NO_DEVICE: return "no usable CUDA device (sm_100a required; there is no CPU fallback)";
    case FEC_ERR_UNSUPPORTED: return "unsupported semantics or format";
    case FEC_ERR_OOM: return "out of memory";
    default: return "unknown error";
  }
}

// contiguous lane range of shard `index` of `n_shards`: boundaries rounded up to multiples of 1024 lanes so that
// every record offset (32/33/64/65 B per lane) stays 16-byte aligned; pure arithmetic, no device needed
void fec_shard_range(size_t n, int n_shards, int index, size_t* first, size_t* count) {
  if (n_shards < 1) n_shards = 1;
  auto cut = [&](int i) -> size_t {
    if (i <= 0) return 0;
    if (i >= n_shards) return n;
    size_t c = (n / (size_t)n_shards * (size_t)i + n % (size_t)n_shards * (size_t)i / (size_t)n_shards + 1023) / 1024 * 1024;
    return c < n ? c : n;
  };
  const size_t a = cut(index), b = cut(index + 1);
  if (first) *first = a;
  if (count) *count = b - a;
}

int fec_ctx_create(const int* device_ids, int n_devices, fec_ctx** out) {
  if (out == nullptr) return FEC_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return FEC_ERR_NO_DEVICE;
  fec_ctx* ctx = new (std::nothrow) fec_ctx();
  if (ctx == nullptr) return FEC_ERR_OOM;
  int cur = 0;
  if (device_ids == nullptr || n_devices <= 0) {
    if (cudaGetDevice(&cur) != cudaSuccess) { delete ctx; return FEC_ERR_NO_DEVICE; }
    ctx->devs.resize(1);
    ctx->devs[0].id = cur;
  } else {
    ctx->devs.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; i++) {
      if (device_ids[i] < 0 || device_ids[i] >= count) { ctx->devs.clear(); delete ctx; return FEC_ERR_INVALID_ARG; }
      ctx->devs[(size_t)i].id = device_ids[i];
    }
  }
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& d : ctx->devs) {
    int rc = device_init(ctx, d);
    if (rc) {
      fprintf(stderr, "forge_ec_cuda: context creation failed: %s (%s)\n", fec_strerror(rc), ctx->last_err);
      for (auto& dd : ctx->devs) device_free(dd);
      delete ctx;
      cudaSetDevice(prev);
      return rc;
    }
  }
  cudaSetDevice(prev);
  *out = ctx;
  return FEC_OK;
}

void fec_ctx_destroy(fec_ctx* ctx) {
  if (ctx == nullptr) return;
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& d : ctx->devs) device_free(d);
  for (cudaEvent_t ev : ctx->stage_ev) cudaEventDestroy(ev);
  cudaSetDevice(prev);
  delete ctx;
}

int fec_ctx_device_count(const fec_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
uint64_t fec_ctx_launch_count(const fec_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
const char* fec_ctx_last_cuda_error(const fec_ctx* ctx) { return ctx ? ctx->last_err : ""; }

int fec_batch_mul(fec_ctx* ctx, fec_curve curve, fec_sem sem, const uint8_t* scalars, const uint8_t* points,
                  fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* out, uint8_t* status) {
  if (ctx == nullptr || (n > 0 && (scalars == nullptr || out == nullptr))) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  if (!sem_supported(sem, curve, points_fmt, (int)out_fmt)) return FEC_ERR_UNSUPPORTED;
  HostJob job{0, (int)curve, (int)sem, (int)points_fmt, (int)out_fmt, 0, scalars, 32, points,
              (size_t)(points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64), out, (size_t)out_size(out_fmt), status};
  return run_host_job(ctx, job, n);
}

int fec_batch_mul_dev(fec_ctx* ctx, int dev_index, fec_curve curve, fec_sem sem, const uint8_t* d_scalars,
                      const uint8_t* d_points, fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* d_out,
                      uint8_t* d_status, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_scalars == nullptr || d_out == nullptr))) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_scalars) || !aligned16(d_points) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  if (!sem_supported(sem, curve, points_fmt, (int)out_fmt)) return FEC_ERR_UNSUPPORTED;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  int rc = launch_mul(ctx, *d, (int)curve, (int)sem, d_scalars, d_points, (int)points_fmt, n, (int)out_fmt, d_out, d_status,
                      (cudaStream_t)stream);
  return rc;
}

// Development / measurement aid (bench.py): while enabled, the secp256k1 k*P path records CUDA events between its
// three kernels for calls made through fec_batch_mul_dev.  fec_ctx_stage_ms waits for them and returns the summed
// milliseconds of {prepare, ladder, finish} and the number of timed chunks, then forgets them.
int fec_ctx_stage_timing(fec_ctx* ctx, int enable) {
  if (ctx == nullptr) return FEC_ERR_INVALID_ARG;
  ctx->stage_timing = enable != 0;
  ctx->stage_used = 0;
  return FEC_OK;
}
int fec_ctx_stage_ms(fec_ctx* ctx, double ms[3], int* chunks) {
  if (ctx == nullptr || ms == nullptr) return FEC_ERR_INVALID_ARG;
  ms[0] = ms[1] = ms[2] = 0.0;
  int c = 0;
  for (int i = 0; i + 3 < ctx->stage_used; i += 4) {
    FEC_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[(size_t)i + 3]));
    for (int s = 0; s < 3; s++) {
      float t = 0.f;
      FEC_CUDA(ctx, cudaEventElapsedTime(&t, ctx->stage_ev[(size_t)i + s], ctx->stage_ev[(size_t)i + s + 1]));
      ms[s] += (double)t;
    }
    c++;
  }
  if (chunks) *chunks = c;
  ctx->stage_used = 0;
  return FEC_OK;
}

static bool ecdh_args_ok(int curve, int pfmt, int kdf, const uint8_t* info, size_t info_len, size_t out_len) {
  if (curve < 0 || curve > 1 || pfmt < 0 || pfmt > 1 || kdf < 0 || kdf > 2) return false;
  if (kdf == KDF_NONE) return true;
  return info_len <= (size_t)KDF_MAX_INFO && out_len >= 1 && out_len <= (size_t)KDF_MAX_OUT && (info != nullptr || info_len == 0);
}
int fec_batch_ecdh(fec_ctx* ctx, fec_curve curve, const uint8_t* scalars, const uint8_t* points, fec_points_fmt points_fmt, size_t n,
                   fec_kdf kdf, const uint8_t* info, size_t info_len, size_t out_len, uint8_t* out, uint8_t* status) {
  if (ctx == nullptr || (n > 0 && (scalars == nullptr || points == nullptr || out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (!ecdh_args_ok((int)curve, (int)points_fmt, (int)kdf, info, info_len, out_len)) return FEC_ERR_INVALID_ARG;
  const size_t rec = kdf == FEC_KDF_NONE ? 32 : out_len;
  HostJob job{3, (int)curve, FEC_SEM_STD, (int)points_fmt, 3, 0, scalars, 32, points,
              (size_t)(points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64), out, rec, status};
  job.kdf = (int)kdf; job.info = info; job.info_len = (int)info_len; job.out_len = (int)out_len;
  return run_host_job(ctx, job, n);
}
int fec_batch_ecdh_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_scalars, const uint8_t* d_points,
                       fec_points_fmt points_fmt, size_t n, fec_kdf kdf, const uint8_t* info, size_t info_len, size_t out_len,
                       uint8_t* d_out, uint8_t* d_status, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_scalars == nullptr || d_points == nullptr || d_out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (!ecdh_args_ok((int)curve, (int)points_fmt, (int)kdf, info, info_len, out_len)) return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_scalars) || !aligned16(d_points) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  if (n == 0) return FEC_OK;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  cudaStream_t st = (cudaStream_t)stream;
  AsyncBuf tmp;  // shared secrets (kdf != NONE) and lane status when the caller wants none; given back to the pool on return
  tmp.st = st;
  const size_t xbytes = kdf == FEC_KDF_NONE ? 0 : n * 32, sbytes = d_status == nullptr ? (n + 15) / 16 * 16 : 0;
  if (xbytes + sbytes > 0) FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&tmp.p), xbytes + sbytes, st));
  uint8_t* dx = tmp.p;
  uint8_t* dst = d_status != nullptr ? d_status : tmp.p + xbytes;
  return launch_ecdh(ctx, *d, (int)curve, d_scalars, d_points, (int)points_fmt, n, (int)kdf, info, (int)info_len, (int)out_len,
                     d_out, dst, dx, st);
}

// sum_i k_i * P_i on one device: products through the batched k*P path (affine X || Y + lane status), then the reduction
static int msm_on_device(fec_ctx* ctx, Device& d, int curve, const uint8_t* ds, const uint8_t* dp, int pfmt, size_t n, int ofmt,
                         uint8_t* dout, uint8_t* dstatus, cudaStream_t st) {
  static const bool no_pippenger = getenv("FORGE_EC_MSM_IMPL") != nullptr && getenv("FORGE_EC_MSM_IMPL")[0] == 'p';  // "products": A/B
  // the bucket method indexes its 17 n (window, lane) entries with 32-bit counters and offsets (k_msm_pippenger.cu): above
  // (2^32 - 1) / 17 = 252 645 135 points the products + reduction path below takes over
  if (n >= (curve == FEC_SECP256K1 ? MSM_PIPPENGER_MIN_K1 : MSM_PIPPENGER_MIN_P256) && n <= MSM_PIPPENGER_MAX && !no_pippenger) {
    AsyncBuf scratch;
    scratch.st = st;
    FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&scratch.p), msm_pippenger_scratch_bytes(n), st));
    int skewed = 0;
    const unsigned limit = (unsigned)(64 * (n >> 15) + 4096);
    FEC_CUDA(ctx, msm_pippenger_launch(curve, st, ds, dp, pfmt, n, scratch.p, ofmt, dout, dstatus, limit, &skewed));
    ctx->launches += skewed ? 3 : 9;
    if (!skewed) return FEC_OK;  // otherwise: pathologically repeated digits, use the products + reduction path below
  }
  const int blocks = msm_blocks(d.sms, n);
  const size_t xy_bytes = n * 64, st_bytes = (n + 15) / 16 * 16, sc_bytes = ((size_t)blocks * 24 + 4) * 4;
  AsyncBuf tmp;
  tmp.st = st;
  FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&tmp.p), xy_bytes + st_bytes + sc_bytes, st));
  int rc = launch_mul(ctx, d, curve, FEC_SEM_STD, ds, dp, pfmt, n, FEC_OUT_AFFINE_XY_BE, tmp.p, tmp.p + xy_bytes, st);
  if (rc) return rc;
  ctx->launches += 2;
  FEC_CUDA(ctx, msm_reduce_launch(curve, d.sms, st, tmp.p, tmp.p + xy_bytes, n, reinterpret_cast<u32*>(tmp.p + xy_bytes + st_bytes), ofmt,
                                  dout, dstatus));
  return FEC_OK;
}
int fec_msm_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_scalars, const uint8_t* d_points,
                fec_points_fmt points_fmt, size_t n, fec_out out_fmt, uint8_t* d_out, uint8_t* d_status, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || n == 0 || d_scalars == nullptr || d_points == nullptr || d_out == nullptr) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_scalars) || !aligned16(d_points)) return FEC_ERR_INVALID_ARG;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  int rc = msm_on_device(ctx, *d, (int)curve, d_scalars, d_points, (int)points_fmt, n, (int)out_fmt, d_out, d_status, (cudaStream_t)stream);
  return rc;
}
// host buffers: every device of the context sums its contiguous lane range, the per-device sums are added on device 0
int fec_msm(fec_ctx* ctx, fec_curve curve, const uint8_t* scalars, const uint8_t* points, fec_points_fmt points_fmt, size_t n,
            fec_out out_fmt, uint8_t* out, uint8_t* status) {
  if (ctx == nullptr || n == 0 || scalars == nullptr || points == nullptr || out == nullptr) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  const size_t G = ctx->devs.size();
  const size_t psz = points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64;
  std::vector<uint8_t> part(G * 64, 0), pst(G, (uint8_t)FEC_LANE_INFINITY);
  std::vector<int> rcs(G, FEC_OK);
  DeviceGuard restore;
  auto shard = [&](size_t i) -> int {
    size_t first = 0, count = 0;
    fec_shard_range(n, (int)G, (int)i, &first, &count);
    if (count == 0) return FEC_OK;
    Device& d = ctx->devs[i];
    FEC_CUDA(ctx, cudaSetDevice(d.id));
    cudaStream_t st = d.streams[0];
    struct Sync { cudaStream_t s; ~Sync() { cudaStreamSynchronize(s); } } sync{st};  // (declared first: runs after the free below)
    AsyncBuf buf;
    buf.st = st;
    const size_t sb = count * 32, pb = (count * psz + 15) / 16 * 16;
    FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&buf.p), sb + pb + 128, st));
    FEC_CUDA(ctx, cudaMemcpyAsync(buf.p, scalars + first * 32, sb, cudaMemcpyHostToDevice, st));
    FEC_CUDA(ctx, cudaMemcpyAsync(buf.p + sb, points + first * psz, count * psz, cudaMemcpyHostToDevice, st));
    int rc = msm_on_device(ctx, d, (int)curve, buf.p, buf.p + sb, (int)points_fmt, count, FEC_OUT_AFFINE_XY_BE, buf.p + sb + pb,
                           buf.p + sb + pb + 64, st);
    if (rc) return rc;
    FEC_CUDA(ctx, cudaMemcpyAsync(&part[i * 64], buf.p + sb + pb, 64, cudaMemcpyDeviceToHost, st));
    FEC_CUDA(ctx, cudaMemcpyAsync(&pst[i], buf.p + sb + pb + 64, 1, cudaMemcpyDeviceToHost, st));
    FEC_CUDA(ctx, cudaStreamSynchronize(st));
    return FEC_OK;
  };
  if (G == 1) {
    rcs[0] = shard(0);
  } else {
    std::vector<std::thread> th;
    for (size_t i = 0; i < G; i++) th.emplace_back([&, i]() { rcs[i] = shard(i); });
    for (auto& t : th) t.join();
  }
  for (int rc : rcs) if (rc) return rc;
  // combine: the G partial sums are G more "products" for the same reduction (status 3 = identity contributes nothing)
  Device& d0 = ctx->devs[0];
  FEC_CUDA(ctx, cudaSetDevice(d0.id));
  cudaStream_t st = d0.streams[0];
  struct Sync { cudaStream_t s; ~Sync() { cudaStreamSynchronize(s); } } sync{st};
  AsyncBuf buf;
  buf.st = st;
  const size_t sc_off = (G * 64 + G + 15) / 16 * 16;
  FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&buf.p), sc_off + (24 + 4) * 4 + 128, st));
  FEC_CUDA(ctx, cudaMemcpyAsync(buf.p, part.data(), G * 64, cudaMemcpyHostToDevice, st));
  FEC_CUDA(ctx, cudaMemcpyAsync(buf.p + G * 64, pst.data(), G, cudaMemcpyHostToDevice, st));
  uint8_t* dres = buf.p + sc_off + (24 + 4) * 4;
  ctx->launches += 2;
  FEC_CUDA(ctx, msm_reduce_launch((int)curve, 1, st, buf.p, buf.p + G * 64, G, reinterpret_cast<u32*>(buf.p + sc_off), (int)out_fmt, dres,
                                  dres + 80));
  uint8_t hst = 0;
  FEC_CUDA(ctx, cudaMemcpyAsync(out, dres, (size_t)out_size(out_fmt), cudaMemcpyDeviceToHost, st));
  FEC_CUDA(ctx, cudaMemcpyAsync(&hst, dres + 80, 1, cudaMemcpyDeviceToHost, st));
  FEC_CUDA(ctx, cudaStreamSynchronize(st));
  if (status) *status = hst;
  return FEC_OK;
}

int fec_batch_point_decode(fec_ctx* ctx, fec_curve curve, const uint8_t* points, fec_points_fmt points_fmt, size_t n, fec_out out_fmt,
                           uint8_t* out, uint8_t* status) {
  if (ctx == nullptr || (n > 0 && (points == nullptr || out == nullptr))) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  // the records travel as "input a" of the pipeline; 33- and 64-byte records use the larger device buffer (rec_a > 32)
  HostJob job{5, (int)curve, FEC_SEM_STD, (int)points_fmt, (int)out_fmt, 0, points, (size_t)(points_fmt == FEC_PT_SEC1_COMPRESSED ? 33 : 64),
              nullptr, 0, out, (size_t)out_size(out_fmt), status};
  return run_host_job(ctx, job, n);
}
int fec_batch_point_decode_dev(fec_ctx* ctx, int dev_index, fec_curve curve, const uint8_t* d_points, fec_points_fmt points_fmt, size_t n,
                               fec_out out_fmt, uint8_t* d_out, uint8_t* d_status, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_points == nullptr || d_out == nullptr))) return FEC_ERR_INVALID_ARG;
  if ((int)curve < 0 || (int)curve > 1 || (int)out_fmt < 0 || (int)out_fmt > 4 || (int)points_fmt < 0 || (int)points_fmt > 1)
    return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_points) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  if (n == 0) return FEC_OK;
  DeviceGuard restore;
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  ctx->launches++;
  FEC_CUDA(ctx, point_decode_launch((int)curve, (cudaStream_t)stream, d_points, (int)points_fmt, n, (int)out_fmt, d_out, d_status));
  return FEC_OK;
}

static bool nonce_args_ok(int curve, int mode) {
  if (mode < 0 || mode > 2) return false;
  if (mode == FEC_NONCE_EDDSA) return true;
  return mode == FEC_NONCE_BIP340 ? curve == FEC_SECP256K1 : (curve == FEC_SECP256K1 || curve == FEC_P256);
}
int fec_batch_nonce_commit(fec_ctx* ctx, fec_curve curve, fec_nonce_mode mode, const uint8_t* nonces, size_t n, uint8_t* out_R,
                           uint8_t* out_r, uint8_t* status) {
  if (ctx == nullptr || (n > 0 && (nonces == nullptr || out_R == nullptr || out_r == nullptr))) return FEC_ERR_INVALID_ARG;
  if (!nonce_args_ok((int)curve, (int)mode)) return FEC_ERR_INVALID_ARG;
  HostJob job{4, (int)curve, FEC_SEM_STD, 0, 0, 0, nonces, (size_t)(mode == FEC_NONCE_EDDSA ? 64 : 32), nullptr, 0, out_R,
              (size_t)(mode == FEC_NONCE_ECDSA ? 33 : 32), status};
  job.mode = (int)mode;
  job.out2 = out_r;
  return run_host_job(ctx, job, n);
}
int fec_batch_nonce_commit_dev(fec_ctx* ctx, int dev_index, fec_curve curve, fec_nonce_mode mode, const uint8_t* d_nonces, size_t n,
                               uint8_t* d_out_R, uint8_t* d_out_r, uint8_t* d_status, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_nonces == nullptr || d_out_R == nullptr || d_out_r == nullptr))) return FEC_ERR_INVALID_ARG;
  if (!nonce_args_ok((int)curve, (int)mode)) return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_nonces) || !aligned16(d_out_R) || !aligned16(d_out_r)) return FEC_ERR_INVALID_ARG;
  if (n == 0) return FEC_OK;
  DeviceGuard restore;
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  cudaStream_t st = (cudaStream_t)stream;
  AsyncBuf tmp;  // affine R and the lane status of the multiplication (Weierstrass modes)
  tmp.st = st;
  const size_t xyb = mode == FEC_NONCE_EDDSA ? 0 : n * 64, stb = mode == FEC_NONCE_EDDSA ? 0 : (n + 15) / 16 * 16;
  if (xyb + stb > 0) FEC_CUDA(ctx, scratch_alloc(reinterpret_cast<void**>(&tmp.p), xyb + stb, st));
  return launch_nonce(ctx, *d, (int)curve, (int)mode, d_nonces, n, d_out_R, d_out_r, d_status, tmp.p, tmp.p + xyb, st);
}

int fec_batch_x25519(fec_ctx* ctx, const uint8_t* scalars, const uint8_t* u, size_t n, uint8_t* out) {
  return fec_batch_x25519_sem(ctx, FEC_SEM_STD, scalars, u, n, out);
}
int fec_batch_x25519_sem(fec_ctx* ctx, fec_sem sem, const uint8_t* scalars, const uint8_t* u, size_t n, uint8_t* out) {
  if (ctx == nullptr || (n > 0 && (scalars == nullptr || out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (sem != FEC_SEM_STD && sem != FEC_SEM_FORGE_QUIRK) return FEC_ERR_INVALID_ARG;
  HostJob job{1, 0, (int)sem, 0, 0, 0, scalars, 32, u, 32, out, 32, nullptr};
  return run_host_job(ctx, job, n);
}
int fec_batch_x25519_dev(fec_ctx* ctx, int dev_index, const uint8_t* d_scalars, const uint8_t* d_u, size_t n,
                         uint8_t* d_out, void* stream) {
  return fec_batch_x25519_sem_dev(ctx, dev_index, FEC_SEM_STD, d_scalars, d_u, n, d_out, stream);
}
int fec_batch_x25519_sem_dev(fec_ctx* ctx, int dev_index, fec_sem sem, const uint8_t* d_scalars, const uint8_t* d_u, size_t n,
                             uint8_t* d_out, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_scalars == nullptr || d_out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (sem != FEC_SEM_STD && sem != FEC_SEM_FORGE_QUIRK) return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_scalars) || !aligned16(d_u) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  int rc = launch_x25519(ctx, *d, (int)sem, d_scalars, d_u, n, d_out, (cudaStream_t)stream);
  return rc;
}

int fec_batch_ed25519_pubkey(fec_ctx* ctx, const uint8_t* seeds_or_scalars, int input_is_seed, size_t n, uint8_t* out) {
  return fec_batch_ed25519_pubkey_sem(ctx, FEC_SEM_STD, seeds_or_scalars, input_is_seed, n, out);
}
// FEC_SEM_FORGE_QUIRK replays derive_public_key, which only exists for seeds
static bool ed_sem_ok(int sem, int input_is_seed) { return sem == FEC_SEM_STD || (sem == FEC_SEM_FORGE_QUIRK && input_is_seed != 0); }
int fec_batch_ed25519_pubkey_sem(fec_ctx* ctx, fec_sem sem, const uint8_t* seeds_or_scalars, int input_is_seed, size_t n, uint8_t* out) {
  if (ctx == nullptr || (n > 0 && (seeds_or_scalars == nullptr || out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (sem != FEC_SEM_STD && sem != FEC_SEM_FORGE_QUIRK) return FEC_ERR_INVALID_ARG;
  if (!ed_sem_ok((int)sem, input_is_seed)) return FEC_ERR_UNSUPPORTED;
  HostJob job{2, 0, (int)sem, 0, 0, input_is_seed, seeds_or_scalars, 32, nullptr, 0, out, 32, nullptr};
  return run_host_job(ctx, job, n);
}
int fec_batch_ed25519_pubkey_dev(fec_ctx* ctx, int dev_index, const uint8_t* d_in, int input_is_seed, size_t n,
                                 uint8_t* d_out, void* stream) {
  return fec_batch_ed25519_pubkey_sem_dev(ctx, dev_index, FEC_SEM_STD, d_in, input_is_seed, n, d_out, stream);
}
int fec_batch_ed25519_pubkey_sem_dev(fec_ctx* ctx, int dev_index, fec_sem sem, const uint8_t* d_in, int input_is_seed, size_t n,
                                     uint8_t* d_out, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && (d_in == nullptr || d_out == nullptr))) return FEC_ERR_INVALID_ARG;
  if (sem != FEC_SEM_STD && sem != FEC_SEM_FORGE_QUIRK) return FEC_ERR_INVALID_ARG;
  if (!ed_sem_ok((int)sem, input_is_seed)) return FEC_ERR_UNSUPPORTED;
  if (!aligned16(d_in) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  int rc = launch_ed(ctx, *d, (int)sem, d_in, input_is_seed, n, d_out, (cudaStream_t)stream);
  return rc;
}

int fec_field_bench_dev(fec_ctx* ctx, int dev_index, fec_prime prime, fec_field_op op, size_t lanes, int iters,
                        const uint8_t* d_a, const uint8_t* d_b, uint8_t* d_out, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (lanes > 0 && (d_a == nullptr || d_b == nullptr || d_out == nullptr)) || iters < 0)
    return FEC_ERR_INVALID_ARG;
  if ((int)prime < 0 || (int)prime > 2 || (int)op < 0 || (int)op > 2) return FEC_ERR_INVALID_ARG;
  if (!aligned16(d_a) || !aligned16(d_b) || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  int rc = launch_field(ctx, *d, (int)prime, (int)op, lanes, iters, d_a, d_b, d_out, (cudaStream_t)stream);
  return rc;
}

int fec_field_bench(fec_ctx* ctx, fec_prime prime, fec_field_op op, size_t lanes, int iters, const uint8_t* a,
                    const uint8_t* b, uint8_t* out, double* seconds) {
  if (ctx == nullptr || (lanes > 0 && (a == nullptr || b == nullptr || out == nullptr)) || iters < 0) return FEC_ERR_INVALID_ARG;
  if ((int)prime < 0 || (int)prime > 2 || (int)op < 0 || (int)op > 2) return FEC_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  Device& d = ctx->devs[0];
  int prev = 0;
  cudaGetDevice(&prev);
  FEC_CUDA(ctx, cudaSetDevice(d.id));
  uint8_t *da = nullptr, *db = nullptr, *dout = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = FEC_OK;
  auto cleanup = [&]() {
    cudaFree(da); cudaFree(db); cudaFree(dout);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaSetDevice(prev);
  };
#define FEC_TRY(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      set_err(ctx, __FILE__, __LINE__, #call, cudaGetErrorString(e_));                                    \
      cleanup();                                                                                        \
      return e_ == cudaErrorMemoryAllocation ? FEC_ERR_OOM : FEC_ERR_CUDA;                              \
    }                                                                                                   \
  } while (0)
  if (lanes > 0) {
    FEC_TRY(cudaMalloc(&da, lanes * 32));
    FEC_TRY(cudaMalloc(&db, lanes * 32));
    FEC_TRY(cudaMalloc(&dout, lanes * 32));
    FEC_TRY(cudaEventCreate(&e0));
    FEC_TRY(cudaEventCreate(&e1));
    cudaStream_t st = d.streams[0];
    FEC_TRY(cudaMemcpyAsync(da, a, lanes * 32, cudaMemcpyHostToDevice, st));
    FEC_TRY(cudaMemcpyAsync(db, b, lanes * 32, cudaMemcpyHostToDevice, st));
    FEC_TRY(cudaEventRecord(e0, st));
    rc = launch_field(ctx, d, (int)prime, (int)op, lanes, iters, da, db, dout, st);
    if (rc) { cleanup(); return rc; }
    FEC_TRY(cudaEventRecord(e1, st));
    FEC_TRY(cudaMemcpyAsync(out, dout, lanes * 32, cudaMemcpyDeviceToHost, st));
    FEC_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    FEC_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (seconds) *seconds = (double)ms * 1e-3;
  } else if (seconds) {
    *seconds = 0.0;
  }
  cleanup();
  return rc;
}

int fec_imad_peak(fec_ctx* ctx, int dev_index, int variant, double* instr_per_sec) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || instr_per_sec == nullptr || variant < 0 || variant > 2) return FEC_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  const int blocks = d->sms * 8, trips = 4096;
  u32* sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(sink);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
  FEC_TRY(cudaMalloc(&sink, (size_t)blocks * PEAK_THREADS * 4));
  FEC_TRY(cudaEventCreate(&e0));
  FEC_TRY(cudaEventCreate(&e1));
  cudaStream_t st = d->streams[0];
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {  // first repetition is the warm-up
    FEC_TRY(cudaEventRecord(e0, st));
    ctx->launches++;
    FEC_TRY(imad_peak_launch(variant, blocks, st, sink, 0x1234567u + rep, trips));
    FEC_TRY(cudaEventRecord(e1, st));
    FEC_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    FEC_TRY(cudaEventElapsedTime(&ms, e0, e1));
    // multiply-add instructions issued: variant 1 issues two per chain step
    const double per_step = (variant == 1) ? 2.0 : 1.0;
    const double instr = (double)blocks * PEAK_THREADS * (double)trips * PEAK_INNER * PEAK_CHAINS * per_step;
    const double rate = instr / ((double)ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  *instr_per_sec = best;
  cleanup();
  return FEC_OK;
}

int fec_synth_dev(fec_ctx* ctx, int dev_index, int mode, uint32_t stream_id, uint32_t sub, uint64_t first_lane, int edge,
                  size_t n, uint8_t* d_out, void* stream) {
  Device* d = dev_at(ctx, dev_index);
  if (d == nullptr || (n > 0 && d_out == nullptr) || mode < 0 || mode > 2 || !aligned16(d_out)) return FEC_ERR_INVALID_ARG;
  if (n == 0) return FEC_OK;
  DeviceGuard restore;  // the caller's current device comes back on every return path
  FEC_CUDA(ctx, cudaSetDevice(d->id));
  const unsigned long long seed = 0x666F7267652D6563ULL;  // "forge-ec"
  ctx->launches++;
  FEC_CUDA(ctx, synth_launch(d->sms, (cudaStream_t)stream, d_out, n, (u32)seed, (u32)(seed >> 32), stream_id, sub,
                             (unsigned long long)first_lane, mode, edge));
  return FEC_OK;
}

}  // extern "C"
