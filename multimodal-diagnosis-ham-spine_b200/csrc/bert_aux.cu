// BERT embedding front end (gather word + position + token-type rows) and its gradient scatter.
// The LayerNorm / dropout that follow use norm.cu.  Tables and their gradients are fp32.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;

namespace {

// e[row, :] = word[ids[row]] + pos[row % S] + type[type_ids ? type_ids[row] : 0]   (fp32, 4 floats/thread)
__global__ void __launch_bounds__(256) embed_gather_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                                                           const float* __restrict__ word, const float* __restrict__ pos,
                                                           const float* __restrict__ type, float* __restrict__ e, int rows,
                                                           int S, int C, int vocab) {
  const int cvec = C >> 2;
  const int64_t total = (int64_t)rows * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const int row = (int)(i / cvec);
    int64_t id = ids[row];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int64_t tt = type_ids ? type_ids[row] : 0;
    const float4 a = *reinterpret_cast<const float4*>(word + id * C + cv * 4);
    const float4 b = *reinterpret_cast<const float4*>(pos + (int64_t)(row % S) * C + cv * 4);
    const float4 c = *reinterpret_cast<const float4*>(type + tt * C + cv * 4);
    *reinterpret_cast<float4*>(e + (int64_t)row * C + cv * 4) =
        make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
  }
}

// Scatter de [rows, C] (fp32) into the three tables' gradients.  One CTA owns `rows_per_block`
// consecutive rows; word rows go out as atomics (ids repeat), position rows are summed over the batch by
// atomics (S*C addresses), the token-type-0 row is first reduced inside the CTA.
__global__ void __launch_bounds__(256) embed_scatter_kernel(const float* __restrict__ de, const int64_t* __restrict__ ids,
                                                            const int64_t* __restrict__ type_ids, float* __restrict__ gword,
                                                            float* __restrict__ gpos, float* __restrict__ gtype, int rows, int S,
                                                            int C, int vocab, int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t0 = 0.f;
    for (int row = r0; row < r1; row++) {
      const float g = de[(int64_t)row * C + c];
      int64_t id = ids[row];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
      if (gword) atomicAdd(gword + id * C + c, g);
      if (gpos) atomicAdd(gpos + (int64_t)(row % S) * C + c, g);
      if (gtype) {
        if (type_ids && type_ids[row] != 0) atomicAdd(gtype + type_ids[row] * C + c, g);
        else t0 += g;
      }
    }
    if (gtype) atomicAdd(gtype + c, t0);
  }
}

}  // namespace

extern "C" int mdhs_embed_gather(const int64_t* ids, const int64_t* type_ids, const float* word, const float* pos,
                                 const float* type, float* e, int rows, int S, int C, int vocab, void* stream) {
  if (!ids || !word || !pos || !type || !e || rows <= 0 || (C % 4)) return MDHS_ERR_ARG;
  int64_t total = (int64_t)rows * (C / 4);
  int grid = (int)((total + 255) / 256);
  if (grid > (int64_t)mdhs_num_sms() * 16) grid = (int64_t)mdhs_num_sms() * 16;
  g_mdhs_launches++;
  embed_gather_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ids, type_ids, word, pos, type, e, rows, S, C, vocab);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_embed_scatter(const float* de, const int64_t* ids, const int64_t* type_ids, float* gword, float* gpos,
                                  float* gtype, int rows, int S, int C, int vocab, void* stream) {
  if (!de || !ids || rows <= 0) return MDHS_ERR_ARG;
  const int rpb = 16;
  g_mdhs_launches++;
  embed_scatter_kernel<<<ceil_div(rows, rpb), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(de, ids, type_ids, gword, gpos, gtype,
                                                                                                rows, S, C, vocab, rpb);
  MDHS_RETURN_LAST();
}
