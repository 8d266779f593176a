// Residue-frame voxeliser: what aposteriori's make-frame-dataset does before the network runs
// (/root/reference/README.md:84-97,242; /root/reference/ui.py:63-87; schema /root/reference/design_utils/utils.py:238-251).
//
// One CTA per residue frame.  Every atom of the structure is moved into the residue's local frame (CA at the origin, N on
// the +y axis, C in the xy plane at x > 0 -- the convention under which the reference's hard-coded C-beta
// (-0.741287356, -0.53937931, -1.224287356) is the mean C-beta of real residues, checked to 0.03 A on 1ubq), dropped if
// its voxel lies outside the V^3 grid, and written into its channel: a single voxel (boolean frames) or a 3x3x3
// gaussian stamp centred on the atom's sub-voxel position and normalised to unit mass (gaussian frames).  Stamps are
// accumulated as 2^-24 fixed-point integers with integer atomics, so the result does not depend on the order in which
// atoms arrive (a float atomicAdd would); voxelise_finalize_kernel turns the integers into the frame dtype.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace tb {

constexpr float kVoxFixed = 16777216.0f;      // 2^24

struct VoxeliseParams {
    const float4* atoms;          // (n_atoms): x, y, z, gaussian sigma in voxel units
    const int32_t* atom_channel;  // (n_atoms): channel of the atom type, < 0: not encoded
    const int32_t* atom_residue;  // (n_atoms): index of the residue the atom belongs to
    const int32_t* atom_is_cb;    // (n_atoms): 1 for C-beta atoms
    int64_t n_atoms;
    const float* res_frame;       // (n_res_total, 12): origin (3), then the rows x, y, z of the rotation
    const float* res_property;    // (n_res_total) value written to the property channel at C-beta positions, or NULL
    const int32_t* res_index;     // (>= res_first + n_res) residues to voxelise, in output order; NULL = identity
    const int32_t* res_atom_range;// (n_res_total, 2) [first, end) of the atoms a residue's frame has to look at (its own
                                  // structure, when several structures share one atom table); NULL = all atoms
    int64_t res_first;
    int32_t V;                    // voxels per side (odd)
    float inv_edge;               // 1 / voxel edge length
    int32_t C;                    // channels
    int32_t gaussian;
    int32_t encode_cb;            // replace the centre residue's C-beta by the ideal one
    float cb_x, cb_y, cb_z, cb_sigma;
    int32_t cb_channel;
    int32_t property_channel;     // < 0: none
    int32_t* scratch;             // (n_res, V, V, V, C) zero-initialised
};

#if defined(__CUDACC__)

// Local coordinates and weights are evaluated in double precision: the CPU oracle then agrees to the last fixed-point bit
// (an fp32 transform with fused multiply-adds would not), and the work per frame is tiny.
__device__ __forceinline__ void voxelise_atom(const VoxeliseParams& p, int32_t* frame, double lx, double ly, double lz,
                                              float sigma, int ch, float prop, bool with_prop) {
    const int half = p.V / 2;
    const double ux = lx * static_cast<double>(p.inv_edge), uy = ly * static_cast<double>(p.inv_edge),
                 uz = lz * static_cast<double>(p.inv_edge);
    const int ix = __double2int_rn(ux) + half, iy = __double2int_rn(uy) + half, iz = __double2int_rn(uz) + half;
    if (ix < 0 || iy < 0 || iz < 0 || ix >= p.V || iy >= p.V || iz >= p.V) return;
    if (!p.gaussian) {
        frame[((static_cast<int64_t>(ix) * p.V + iy) * p.V + iz) * p.C + ch] = static_cast<int32_t>(kVoxFixed);
        if (with_prop) atomicAdd(&frame[((static_cast<int64_t>(ix) * p.V + iy) * p.V + iz) * p.C + p.property_channel],
                                 __float2int_rn(prop * kVoxFixed));
        return;
    }
    const double dx = ux - (ix - half), dy = uy - (iy - half), dz = uz - (iz - half);
    const double inv2s2 = 1.0 / (2.0 * static_cast<double>(sigma) * static_cast<double>(sigma));
    double wx[3], wy[3], wz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        wx[k] = exp(-((k - 1) - dx) * ((k - 1) - dx) * inv2s2);
        wy[k] = exp(-((k - 1) - dy) * ((k - 1) - dy) * inv2s2);
        wz[k] = exp(-((k - 1) - dz) * ((k - 1) - dz) * inv2s2);
    }
    const double norm = 1.0 / ((wx[0] + wx[1] + wx[2]) * (wy[0] + wy[1] + wy[2]) * (wz[0] + wz[1] + wz[2]));
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int x = ix + a - 1, y = iy + b - 1, z = iz + c - 1;
                if (x < 0 || y < 0 || z < 0 || x >= p.V || y >= p.V || z >= p.V) continue;
                const double w = wx[a] * wy[b] * wz[c] * norm;
                const int64_t o = ((static_cast<int64_t>(x) * p.V + y) * p.V + z) * p.C;
                atomicAdd(&frame[o + ch], __double2int_rn(w * static_cast<double>(kVoxFixed)));
                if (with_prop) atomicAdd(&frame[o + p.property_channel], __double2int_rn(w * static_cast<double>(prop) * static_cast<double>(kVoxFixed)));
            }
}

__global__ void voxelise_kernel(const VoxeliseParams p) {
    const int64_t r = p.res_index ? p.res_index[p.res_first + blockIdx.x] : p.res_first + blockIdx.x;
    const float* f = p.res_frame + r * 12;
    const float ox = f[0], oy = f[1], oz = f[2];
    int32_t* frame = p.scratch + static_cast<int64_t>(blockIdx.x) * p.V * p.V * p.V * p.C;
    const float prop_self = p.res_property ? p.res_property[r] : 0.f;
    const int64_t a_first = p.res_atom_range ? p.res_atom_range[2 * r] : 0;
    const int64_t a_end = p.res_atom_range ? p.res_atom_range[2 * r + 1] : p.n_atoms;
    for (int64_t a = a_first + threadIdx.x; a < a_end; a += blockDim.x) {
        const int ch = p.atom_channel[a];
        if (ch < 0) continue;
        const bool is_cb = p.atom_is_cb[a] != 0;
        if (p.encode_cb && is_cb && p.atom_residue[a] == r) continue;       // replaced by the ideal C-beta below
        const float4 q = p.atoms[a];
        const double tx = static_cast<double>(q.x) - ox, ty = static_cast<double>(q.y) - oy, tz = static_cast<double>(q.z) - oz;
        const double lx = __dadd_rn(__dadd_rn(__dmul_rn(f[3], tx), __dmul_rn(f[4], ty)), __dmul_rn(f[5], tz));
        const double ly = __dadd_rn(__dadd_rn(__dmul_rn(f[6], tx), __dmul_rn(f[7], ty)), __dmul_rn(f[8], tz));
        const double lz = __dadd_rn(__dadd_rn(__dmul_rn(f[9], tx), __dmul_rn(f[10], ty)), __dmul_rn(f[11], tz));
        const bool with_prop = p.property_channel >= 0 && is_cb && p.res_property != nullptr;
        voxelise_atom(p, frame, lx, ly, lz, q.w, ch, with_prop ? p.res_property[p.atom_residue[a]] : 0.f, with_prop);
    }
    if (p.encode_cb && threadIdx.x == 0)
        voxelise_atom(p, frame, p.cb_x, p.cb_y, p.cb_z, p.cb_sigma, p.cb_channel, prop_self,
                      p.property_channel >= 0 && p.res_property != nullptr);
}

// fixed point -> frame dtype (float32 / float16; uint8 for boolean frames: any non-zero count is 1)
template <typename T>
__global__ void voxelise_finalize_kernel(const int32_t* __restrict__ scratch, int64_t n, int boolean, T* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = boolean ? (scratch[i] != 0 ? 1.0f : 0.0f) : static_cast<float>(scratch[i]) * (1.0f / kVoxFixed);
        out[i] = static_cast<T>(v);
    }
}

#endif  // __CUDACC__

}  // namespace tb
