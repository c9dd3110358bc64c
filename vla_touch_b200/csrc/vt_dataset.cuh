// Minibatch assembly from the HBM-resident episode store.
//
// Reference: ControllerDataset.__getitem__ (controller_dataset.py:101-170) + the default collate of its DataLoader
// (controller_dataset.py:467-476) + DiffusionControllerTrainer._prepare_batch_for_diffusion (bridge_train.py:120-145):
// per sample (episode, start frame) the reference opens an HDF5 file, slices ctx+horizon frames of every low-dimensional
// stream, divides the gripper column of the ACTION rows by 255, converts to fp32 tensors, collates B samples, uploads them
// and normalises the two action chunks.  Here every stream of every episode sits concatenated in HBM (episode_store.py), a
// sample is four contiguous segments of those arrays, and ONE launch writes the whole collated batch -- states, expert /
// VLA chunks (raw and normalised), forces, marker displacements -- plus the cached DinoV2 features of the last context frame
// of both cameras, chosen between the two cached normalisation branches by the reference's batch-global predicate
// `images.mean() < 0.5` (visual_encoder.py:95-106), evaluated here from per-frame means.
//
// HBM-bound byte moving: ~40 KB per sample, coalesced 4-byte accesses, one CTA per sample.  The values copied are bit-identical
// to the reference's tensors: the two /255 columns are pre-divided on the host in the source precision when the store is built
// (the reference divides in float64 / float32 BEFORE the fp32 cast), the normalisation repeats affine_kernel's operation order.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vt_b200.h"

namespace vt {

// controller_dataset.py:303-345, identical operation order to affine_kernel (vt_elem.cuh): no FMA contraction
__device__ __forceinline__ float normalize_one(float v, float lo, float hi, float pad) {
  const float orig = __fsub_rn(hi, lo);
  const float padded = __fmul_rn(orig, pad);
  const float center = __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
  const float half = __fdiv_rn(padded, 2.0f);
  const float pmin = __fsub_rn(center, half);
  const float pmax = __fadd_rn(center, half);
  float range = __fsub_rn(pmax, pmin);
  if (range < 1e-6f) range = 1.0f;
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(v, pmin)), range), 1.0f);
}

__global__ void __launch_bounds__(256) batch_gather_kernel(const vt_batch_gather_desc d) {
  __shared__ double red[2][256];
  __shared__ int s_branch[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ctx = d.context_frames, H = d.horizon, L = ctx + H, A = d.A;
  const long long f0 = d.start[b];

  // ---- batch-global predicate per camera (every CTA evaluates it in the same fixed order => same answer everywhere) ----
  if (d.feats) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = tid; i < d.B; i += 256) {
      const long long f = d.start[i] + ctx - 1;
      s0 += d.frame_mean[2 * f];
      s1 += d.frame_mean[2 * f + 1];
    }
    red[0][tid] = s0;
    red[1][tid] = s1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (tid < w) {
        red[0][tid] += red[0][tid + w];
        red[1][tid] += red[1][tid + w];
      }
      __syncthreads();
    }
    if (tid < 2) {
      const int br = (red[tid][0] / (double)d.B) < 0.5 ? 0 : 1;   // mean < 0.5: the images go in un-normalised
      s_branch[tid] = br;
      if (b == 0 && d.branch) d.branch[tid] = br;
    }
    __syncthreads();
    const long long f = f0 + ctx - 1;
    const float* src1 = d.feats + ((f * 2 + 0) * 2 + s_branch[0]) * (long long)d.D;
    const float* src2 = d.feats + ((f * 2 + 1) * 2 + s_branch[1]) * (long long)d.D;
    for (int i = tid; i < d.D; i += 256) {
      d.feat_cam1[(long long)b * d.D + i] = src1[i];
      d.feat_cam2[(long long)b * d.D + i] = src2[i];
    }
  }

  // ---- states [L][A]: context rows as stored, action rows with the gripper column / 255 (controller_dataset.py:120-124:
  //      `future_states` is a VIEW of qpos, so the in-place division shows in `states` too) ----
  const float* q = d.qpos + f0 * A;
  const bool norm = d.action_mins != nullptr;
  for (int i = tid; i < L * A; i += 256) {
    const int r = i / A, a = i - r * A;
    float v = q[i];
    if (r >= ctx && a == A - 1) v = d.grip_scaled[f0 + r];
    if (d.states) d.states[(long long)b * L * A + i] = v;
    if (r >= ctx) {
      const long long o = (long long)b * H * A + (i - ctx * A);
      if (d.expert_actions) d.expert_actions[o] = v;
      if (norm && d.expert_n) d.expert_n[o] = normalize_one(v, d.action_mins[a], d.action_maxs[a], d.pad);
    }
  }
  // ---- VLA chunk predicted at the first action frame, first H rows (controller_dataset.py:128-130) ----
  {
    const long long fc = f0 + ctx;
    const float* v0 = d.vla + fc * (long long)d.vla_T * A;
    for (int i = tid; i < H * A; i += 256) {
      const int r = i / A, a = i - r * A;
      float v = v0[i];
      if (a == A - 1) v = d.vla_last_scaled[fc * d.vla_T + r];
      const long long o = (long long)b * H * A + i;
      if (d.vla_actions) d.vla_actions[o] = v;
      if (norm && d.vla_n) d.vla_n[o] = normalize_one(v, d.vla_mins[a], d.vla_maxs[a], d.pad);
    }
  }
  // ---- forces / marker displacements of all L frames (controller_dataset.py:132-133) ----
  if (d.forces_out) {
    const float* s = d.forces + f0 * d.Fd;
    for (int i = tid; i < L * d.Fd; i += 256) d.forces_out[(long long)b * L * d.Fd + i] = s[i];
  }
  if (d.disps_out && d.disps) {
    const float* s = d.disps + f0 * d.Dd;
    float* o = d.disps_out + (long long)b * L * d.Dd;
    const int n = L * d.Dd;
    if (((d.Dd & 1) == 0) && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(o)) & 7) == 0) {
      const float2* s2 = reinterpret_cast<const float2*>(s);
      float2* o2 = reinterpret_cast<float2*>(o);
      for (int i = tid; i < n / 2; i += 256) o2[i] = s2[i];
    } else {
      for (int i = tid; i < n; i += 256) o[i] = s[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// RDT -> controller hand-off.  Reference, four tensor ops and a host round trip apart
// (scripts/franka_model_eef.py:199-222,312: `action[:, :, AGILEX_STATE_INDICES] * [1,..,1,255]` in the policy's dtype, `.to(float32)`;
// scripts/franka_inference_eef.py:186,546,552-554: `.cpu().numpy()` copy for the buffer, `vla_tensor[:, :, -1] /= 255`, slice
// `[:, :act_chunk_execute_step]` into controller.predict): one launch on the policy's own stream reads the unified action vector
// where the policy left it and writes both fp32 tensors.  Same roundings as the reference: the product is rounded to the policy's
// dtype (bf16 RN) before the fp32 widening, the division is an IEEE fp32 division.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_handoff_kernel(const void* __restrict__ action, int dtype, int B, int N, int S,
                                                            const int32_t* __restrict__ idx, const float* __restrict__ scale, int A,
                                                            float last_div, float* __restrict__ raw, float* __restrict__ chunk, int T_exec) {
  const long long total = (long long)B * N * A;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % A);
    const long long bn = i / A;
    const int n = (int)(bn % N);
    const long long b = bn / N;
    const long long src = bn * S + idx[a];
    float v;
    if (dtype == VT_BF16) {
      const float x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(action)[src]);
      v = __bfloat162float(__float2bfloat16_rn(__fmul_rn(x, scale[a])));     // bf16 * bf16 -> bf16 (torch computes in fp32, rounds once)
    } else {
      v = __fmul_rn(reinterpret_cast<const float*>(action)[src], scale[a]);
    }
    if (raw) raw[i] = v;
    if (chunk && n < T_exec) chunk[(b * T_exec + n) * A + a] = (a == A - 1) ? __fdiv_rn(v, last_div) : v;
  }
}

}  // namespace vt
