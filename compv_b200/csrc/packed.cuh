// Packed arithmetic of sm_100a used by the TMA fast paths (canny_fast.cuh, convlt_fast.cuh).
#pragma once

namespace cvb {

// ---- packed arithmetic (sm_100a): two IEEE fp32 lanes per instruction (FFMA2 / FADD2 / FMUL2: each lane rounds exactly like the scalar instruction, so the
// reference's fma chain is reproduced bit for bit at half the issue slots), and two 16-bit lanes in plain 32-bit integer adds (all intermediate values are kept non-negative by a bias).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(unsigned int lo, unsigned int hi)
{
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
	return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, unsigned int& lo, unsigned int& hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c)
{
	f32x2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 fadd2_rz(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
// bytes ia of wa and ib of wb as the float pair (wa.byte[ia], wb.byte[ib])
__device__ __forceinline__ f32x2 u8x2_to_f32x2(unsigned int wa, int ia, unsigned int wb, int ib, f32x2 negMagic)
{
	return fadd2(pk2(__byte_perm(wa, 0x4B000000u, 0x7440u | ia), __byte_perm(wb, 0x4B000000u, 0x7440u | ib)), negMagic);
}

} // namespace cvb
