// tape.cu -- whole-network evaluation from a non-Python host: dsk_plan_* / dsk_denoiser_fwd.
//
// The network's launch list is assembled in Python (models/nets/punetg.py: the per-(batch, shape, precision) plan).  A TAPE is
// that list recorded once (diffsci_b200/tape.py): for every entry point of include/diffsci_b200.h that was called, its name and
// its arguments as {integer, float, pointer = (buffer, offset), descriptor by value, external tensor, stream}; the sizes of all
// device buffers; the contents of the constant ones (packed weights, parameters, pointer tables and the relocations inside
// them).  This file parses a tape, lays the buffers out inside ONE caller-provided workspace, uploads the constants and replays
// the launches -- the call table is generated from the header (tools/gen_tape_dispatch.py -> tape_dispatch.inc), so every
// `int dsk_*` entry point is replayable.  Replaces KarrasModule.get_denoiser (karrasmodule.py:673-719) for C / C++ hosts.
//
// Tape layout (little-endian, every section 8-byte aligned):
//   header   : "DSKTAPE1", u32 version, n_buffers, n_ops, n_relocs, u64 blob_bytes, u64 content_bytes,
//              i32 batch, i32 channels, i64 sample_elems
//   buffers  : n_buffers x {u64 nbytes, u64 content_offset (~0: no initial contents)}
//   ops      : n_ops x {char name[48], u32 nargs, u32 pad, nargs x {u32 kind, u32 buf, u64 value}}
//   relocs   : n_relocs x {u32 buf, u32 target_buf, u64 offset, u64 target_offset}   (a device pointer stored INSIDE a constant
//              buffer, e.g. the pointer tables of the grouped time-MLP kernels)
//   blob     : descriptor structs passed by pointer (dsk_conv_desc)
//   contents : initial bytes of the constant buffers
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace {

union Slot {
  void* p;
  int64_t i;
  double f;
};
struct FnRow {
  const char* name;
  int nargs;
  const char* kinds;
};
#define TAPE_TABLE
#define TAPE_FN(idx, name, nargs, kinds) {name, nargs, kinds},
const FnRow kFns[] = {
#include "tape_dispatch.inc"
};
#undef TAPE_FN
#undef TAPE_TABLE
constexpr int kNumFns = (int)(sizeof(kFns) / sizeof(kFns[0]));

int call_fn(int id, const Slot* a) {
  switch (id) {
#define TAPE_CALLS
#include "tape_dispatch.inc"
#undef TAPE_CALLS
    default: break;
  }
  return DSK_ERR_ARG;
}

enum ArgKind : uint32_t { A_INT = 0, A_FLT = 1, A_BUF = 2, A_NULL = 3, A_BLOB = 4, A_EXT = 5, A_STREAM = 6 };
enum ExtSlot : uint32_t { EXT_X = 0, EXT_SIGMA = 1, EXT_OUT = 2, EXT_COUNT = 3 };

struct TapeHeader {
  char magic[8];
  uint32_t version, n_buffers, n_ops, n_relocs;
  uint64_t blob_bytes, content_bytes;
  int32_t batch, channels;
  int64_t sample_elems;
};
struct TapeBuffer { uint64_t nbytes, content_off; };
struct TapeArg { uint32_t kind, buf; uint64_t value; };
struct TapeReloc { uint32_t buf, target_buf; uint64_t off, target_off; };
struct Op {
  int fn;
  std::vector<TapeArg> args;
};

}  // namespace

struct dsk_plan {
  TapeHeader h;
  std::vector<TapeBuffer> buffers;
  std::vector<uint64_t> offset;          // of each buffer inside the workspace (256-byte aligned)
  std::vector<Op> ops;
  std::vector<TapeReloc> relocs;
  std::vector<uint8_t> blob, content;
  uint64_t workspace_bytes = 0;
  uint8_t* ws = nullptr;                  // bound workspace
};

using namespace dsk;

extern "C" int dsk_plan_create_from_tape(const void* tape, int64_t nbytes, dsk_plan** out) {
  DSK_REQUIRE(tape && out && nbytes >= (int64_t)sizeof(TapeHeader), "dsk_plan_create_from_tape: bad arguments");
  const uint8_t* p = (const uint8_t*)tape;
  const uint8_t* end = p + nbytes;
  dsk_plan* pl = new dsk_plan();
  auto fail = [&](const char* why) { set_error("dsk_plan_create_from_tape: %s", why); delete pl; return DSK_ERR_ARG; };
  memcpy(&pl->h, p, sizeof(TapeHeader));
  p += sizeof(TapeHeader);
  if (memcmp(pl->h.magic, "DSKTAPE1", 8) != 0 || pl->h.version != 1) return fail("not a DSKTAPE1 version-1 tape");
  if ((uint64_t)(end - p) < (uint64_t)pl->h.n_buffers * sizeof(TapeBuffer)) return fail("truncated (buffers)");
  pl->buffers.resize(pl->h.n_buffers);
  memcpy(pl->buffers.data(), p, pl->buffers.size() * sizeof(TapeBuffer));
  p += pl->buffers.size() * sizeof(TapeBuffer);
  pl->ops.resize(pl->h.n_ops);
  for (auto& op : pl->ops) {
    if (end - p < 56) return fail("truncated (ops)");
    char name[49];
    memcpy(name, p, 48);
    name[48] = 0;
    uint32_t nargs;
    memcpy(&nargs, p + 48, 4);
    p += 56;
    op.fn = -1;
    for (int k = 0; k < kNumFns; ++k)
      if (strcmp(kFns[k].name, name) == 0) { op.fn = k; break; }
    if (op.fn < 0) { set_error("dsk_plan_create_from_tape: unknown entry point '%s'", name); delete pl; return DSK_ERR_ARG; }
    if ((int)nargs != kFns[op.fn].nargs) { set_error("dsk_plan_create_from_tape: %s recorded with %u arguments, the header declares %d", name, nargs, kFns[op.fn].nargs); delete pl; return DSK_ERR_ARG; }
    if ((uint64_t)(end - p) < (uint64_t)nargs * sizeof(TapeArg)) return fail("truncated (arguments)");
    op.args.resize(nargs);
    memcpy(op.args.data(), p, nargs * sizeof(TapeArg));
    p += nargs * sizeof(TapeArg);
    for (uint32_t j = 0; j < nargs; ++j) {
      const TapeArg& a = op.args[j];
      const char kind = kFns[op.fn].kinds[j];
      const bool ptr_kind = a.kind == A_BUF || a.kind == A_NULL || a.kind == A_BLOB || a.kind == A_EXT || a.kind == A_STREAM;
      if ((kind == 'p') != ptr_kind || (kind == 'f' && a.kind != A_FLT) || (kind == 'i' && a.kind != A_INT)) {
        set_error("dsk_plan_create_from_tape: %s argument %u has kind %u where the header declares '%c'", name, j, a.kind, kind);
        delete pl;
        return DSK_ERR_ARG;
      }
      if (a.kind == A_BUF && (a.buf >= pl->h.n_buffers || a.value > pl->buffers[a.buf].nbytes)) return fail("pointer outside its buffer");
      if (a.kind == A_EXT && a.buf >= EXT_COUNT) return fail("bad external slot");
      if (a.kind == A_BLOB && a.value >= pl->h.blob_bytes) return fail("descriptor outside the blob section");
    }
  }
  if ((uint64_t)(end - p) < (uint64_t)pl->h.n_relocs * sizeof(TapeReloc) + pl->h.blob_bytes + pl->h.content_bytes) return fail("truncated (tail)");
  pl->relocs.resize(pl->h.n_relocs);
  memcpy(pl->relocs.data(), p, pl->relocs.size() * sizeof(TapeReloc));
  p += pl->relocs.size() * sizeof(TapeReloc);
  pl->blob.assign(p, p + pl->h.blob_bytes);
  p += pl->h.blob_bytes;
  pl->content.assign(p, p + pl->h.content_bytes);
  uint64_t off = 0;
  pl->offset.resize(pl->h.n_buffers);
  for (uint32_t b = 0; b < pl->h.n_buffers; ++b) {
    pl->offset[b] = off;
    off += (pl->buffers[b].nbytes + 255) & ~(uint64_t)255;
    const uint64_t co = pl->buffers[b].content_off;
    if (co != ~(uint64_t)0 && co + pl->buffers[b].nbytes > pl->h.content_bytes) return fail("buffer contents outside the content section");
  }
  for (const TapeReloc& r : pl->relocs) {
    if (r.buf >= pl->h.n_buffers || r.target_buf >= pl->h.n_buffers || r.off + 8 > pl->buffers[r.buf].nbytes ||
        pl->buffers[r.buf].content_off == ~(uint64_t)0)
      return fail("bad relocation");
  }
  pl->workspace_bytes = off;
  *out = pl;
  return DSK_OK;
}

extern "C" int dsk_plan_load(const char* path, dsk_plan** out) {
  DSK_REQUIRE(path && out, "dsk_plan_load: null argument");
  FILE* f = fopen(path, "rb");
  DSK_REQUIRE(f != nullptr, "dsk_plan_load: cannot open %s", path);
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> data((size_t)(n > 0 ? n : 0));
  const size_t got = n > 0 ? fread(data.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  DSK_REQUIRE(n > 0 && got == (size_t)n, "dsk_plan_load: cannot read %s", path);
  return dsk_plan_create_from_tape(data.data(), (int64_t)data.size(), out);
}

extern "C" int64_t dsk_plan_info(const dsk_plan* pl, int what) {
  if (pl == nullptr) return -1;
  switch (what) {
    case DSK_PLAN_WORKSPACE_BYTES: return (int64_t)pl->workspace_bytes;
    case DSK_PLAN_BATCH: return pl->h.batch;
    case DSK_PLAN_CHANNELS: return pl->h.channels;
    case DSK_PLAN_SAMPLE_ELEMS: return pl->h.sample_elems;
    case DSK_PLAN_LAUNCHES: return (int64_t)pl->ops.size();
    default: return -1;
  }
}

extern "C" int dsk_plan_bind(dsk_plan* pl, void* workspace, void* stream) {
  DSK_REQUIRE(pl && workspace, "dsk_plan_bind: null argument");
  DSK_REQUIRE(((uintptr_t)workspace & 255) == 0, "dsk_plan_bind: the workspace must be 256-byte aligned");
  pl->ws = (uint8_t*)workspace;
  // device pointers stored inside constant buffers: patch the host copy to this workspace, then upload
  for (const TapeReloc& r : pl->relocs) {
    const uint64_t addr = (uint64_t)(uintptr_t)(pl->ws + pl->offset[r.target_buf] + r.target_off);
    memcpy(pl->content.data() + pl->buffers[r.buf].content_off + r.off, &addr, 8);
  }
  cudaStream_t st = as_stream(stream);
  for (uint32_t b = 0; b < pl->h.n_buffers; ++b) {
    const uint64_t co = pl->buffers[b].content_off;
    if (co == ~(uint64_t)0 || pl->buffers[b].nbytes == 0) continue;
    cudaError_t e = cudaMemcpyAsync(pl->ws + pl->offset[b], pl->content.data() + co, pl->buffers[b].nbytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { set_error("dsk_plan_bind: upload of buffer %u failed: %s", b, cudaGetErrorString(e)); return DSK_ERR_CUDA; }
  }
  return DSK_OK;
}

extern "C" int dsk_denoiser_fwd(dsk_plan* pl, const float* x, const float* sigma, float* out, void* stream) {
  DSK_REQUIRE(pl && x && sigma && out, "dsk_denoiser_fwd: null argument");
  DSK_REQUIRE(pl->ws != nullptr, "dsk_denoiser_fwd: the plan is not bound to a workspace (dsk_plan_bind)");
  void* ext[EXT_COUNT] = {(void*)x, (void*)sigma, (void*)out};
  Slot slots[32];
  for (const Op& op : pl->ops) {
    const size_t n = op.args.size();
    DSK_REQUIRE(n <= 32, "dsk_denoiser_fwd: too many arguments");
    for (size_t j = 0; j < n; ++j) {
      const TapeArg& a = op.args[j];
      switch (a.kind) {
        case A_INT: slots[j].i = (int64_t)a.value; break;
        case A_FLT: memcpy(&slots[j].f, &a.value, 8); break;
        case A_BUF: slots[j].p = pl->ws + pl->offset[a.buf] + a.value; break;
        case A_NULL: slots[j].p = nullptr; break;
        case A_BLOB: slots[j].p = pl->blob.data() + a.value; break;
        case A_EXT: slots[j].p = (uint8_t*)ext[a.buf] + a.value; break;
        default: slots[j].p = stream; break;
      }
    }
    const int rc = call_fn(op.fn, slots);
    if (rc != DSK_OK) return rc;          // the entry point has set the error text
  }
  return DSK_OK;
}

extern "C" int dsk_plan_destroy(dsk_plan* pl) {
  delete pl;
  return DSK_OK;
}

// ---- EDM preconditioner scalars on the device (preconditioners.py:30-53; the Python module evaluates them with torch ops) ------
namespace dsk {
__global__ void edm_coeffs_kernel(const float* __restrict__ sigma, float sd, float* c_in, float* c_out, float* c_skip, float* c_noise, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float s = sigma[b];
  const float var = __fadd_rn(__fmul_rn(s, s), __fmul_rn(sd, sd));      // sigma ** 2 + sigma_data ** 2, no contraction
  const float rt = sqrtf(var);
  if (c_in) c_in[b] = 1.0f / rt;
  if (c_out) c_out[b] = __fmul_rn(s, sd) / rt;
  if (c_skip) c_skip[b] = __fmul_rn(sd, sd) / var;
  if (c_noise) c_noise[b] = 0.5f * logf(s);
}
}  // namespace dsk

extern "C" int dsk_edm_coeffs(const float* sigma, float sigma_data, float* c_in, float* c_out, float* c_skip, float* c_noise, int B,
                              void* stream) {
  DSK_REQUIRE(sigma && B > 0, "dsk_edm_coeffs: bad arguments");
  DSK_LAUNCH(edm_coeffs_kernel, (B + 127) / 128, 128, 0, as_stream(stream), sigma, sigma_data, c_in, c_out, c_skip, c_noise, B);
  return DSK_OK;
}
