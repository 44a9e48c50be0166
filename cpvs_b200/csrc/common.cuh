// Shared device helpers for the cpvs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpvs {

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint16_t u16;

constexpr int kMaxLevels = 32;
constexpr int kScanThreads = 256;  // threads per scan tile
constexpr int kScanItems = 4;      // items per thread
constexpr int kScanTile = kScanThreads * kScanItems;

// Node coordinates packed for the breadth-first work lists: x:20 | y:20 | z:24.
__host__ __device__ inline u64 packCoord(u32 x, u32 y, u32 z) { return (u64)x | ((u64)y << 20) | ((u64)z << 40); }
__host__ __device__ inline void unpackCoord(u64 c, u32& x, u32& y, u32& z) {
	x = (u32)(c & 0xFFFFFu);
	y = (u32)((c >> 20) & 0xFFFFFu);
	z = (u32)(c >> 40);
}

// ---- classification, literally as the reference writes it (no FMA contraction: explicit _rn ops) ----
// cs::visible (reference src/CompressedShadowUtil.h:35-44): 1 visible, 0 shadow, 2 partial.
__device__ __forceinline__ u32 classifyRange(float minZ, float maxZ, float minDepth, float maxDepth) {
	if (maxZ <= minDepth) return 1u;
	if (minZ >= maxDepth) return 0u;
	return 2u;
}
// cs::absoluteVisible (reference src/CompressedShadowUtil.h:51-57).
__device__ __forceinline__ u32 classifyPoint(float minZ, float maxZ, float depth) {
	const float midZ = __fmul_rn(__fadd_rn(minZ, maxZ), 0.5f);
	return (midZ <= depth) ? 1u : 0u;
}
// std::min / std::max as MinMaxHierarchy uses them (reference src/MinMaxHierarchy.cpp:29-33,46-47).
__device__ __forceinline__ float stdMin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float stdMax(float a, float b) { return (a < b) ? b : a; }

// 256-bit read-only load (LDG.E.256, new with sm_100: one lane fetches a whole 32-byte sector) that does not allocate in
// L1 and asks L2 to fetch the whole 128-byte line on a miss. Used where a warp touches isolated 32-byte sectors whose
// neighbours are needed by nearby work items moments later: DRAM then sees full-line bursts instead of scattered sectors.
struct Float8 {
	float v[8];
};
__device__ __forceinline__ Float8 ldSector256(const float* p) {
	Float8 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
				 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
				 : "l"(p));
	return r;
}

__device__ __forceinline__ u64 mix64(u64 h) {
	h ^= h >> 33;
	h *= 0xFF51AFD7ED558CCDull;
	h ^= h >> 33;
	h *= 0xC4CEB9FE1A85EC53ull;
	h ^= h >> 33;
	return h;
}

// ---- single-pass prefix sums over tiles (decoupled look-back) -----------------------------------
// One ScanTileState per tile, two running sums (a, b) per scan. A tile publishes its own aggregate
// (status 1), then, once it knows the sum of everything in front of it, its inclusive prefix (status
// 2). Each 64-bit word carries its status in the top two bits next to the value, so a word is valid on
// its own: plain relaxed 64-bit loads and stores suffice, no fences (release/acquire pairs compile to
// MEMBAR.ALL.GPU / CCTL.IVALL here, far too heavy for a per-tile handshake).
struct ScanTileState {
	u64 a;
	u64 b;
};
constexpr u64 kScanValueMask = (1ull << 62) - 1;

struct ScanLaunch {
	u32* ticket;           // zeroed counter handing out tile indices in start order
	ScanTileState* tiles;  // zeroed, one per tile
};

__device__ __forceinline__ u32 ldRelaxed32(const u32* p) {
	u32 v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stRelaxed32(u32* p, u32 v) {
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ldRelaxed64(const u64* p) {
	u64 v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stRelaxed64(u64* p, u64 v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Tile index in start order, so every tile in front of this one is already running or done.
__device__ __forceinline__ u32 scanAcquireTile(const ScanLaunch& sl) {
	__shared__ u32 sTile;
	if (threadIdx.x == 0) sTile = atomicAdd(sl.ticket, 1u);
	__syncthreads();
	return sTile;
}

// Block-wide exclusive scan of one (a,b) pair per thread. Returns this thread's exclusive prefix
// inside the block in (a,b) and the block totals in (totA,totB). kThreads threads per block.
template <int kThreads = kScanThreads>
__device__ __forceinline__ void blockExclusiveScan2(u64& a, u64& b, u64& totA, u64& totB) {
	__shared__ u64 sA[kThreads / 32], sB[kThreads / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u64 ia = a, ib = b;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const u64 ua = __shfl_up_sync(0xFFFFFFFFu, ia, d), ub = __shfl_up_sync(0xFFFFFFFFu, ib, d);
		if (lane >= d) {
			ia += ua;
			ib += ub;
		}
	}
	if (lane == 31) {
		sA[warp] = ia;
		sB[warp] = ib;
	}
	__syncthreads();
	u64 offA = 0, offB = 0, tA = 0, tB = 0;
#pragma unroll
	for (int w = 0; w < kThreads / 32; ++w) {
		if (w < warp) {
			offA += sA[w];
			offB += sB[w];
		}
		tA += sA[w];
		tB += sB[w];
	}
	__syncthreads();
	a = offA + ia - a;
	b = offB + ib - b;
	totA = tA;
	totB = tB;
}

// Publishes this tile's totals and returns the sum over all tiles in front of it (all threads get it).
__device__ __forceinline__ void scanLookback2(const ScanLaunch& sl, u32 tile, u64 totA, u64 totB, u64& preA, u64& preB) {
	__shared__ u64 sPre[2];
	ScanTileState* st = sl.tiles;
	if (threadIdx.x < 32) {
		const int lane = threadIdx.x;
		if (lane == 0) {
			const u64 status = tile == 0 ? (2ull << 62) : (1ull << 62);
			stRelaxed64(&st[tile].a, status | totA);
			stRelaxed64(&st[tile].b, status | totB);
		}
		u64 runA = 0, runB = 0;
		if (tile > 0) {
			long long look = (long long)tile - 1;
			for (;;) {
				const long long idx = look - lane;
				u64 wa = 2ull << 62, wb = 2ull << 62;  // lanes in front of tile 0 report "inclusive, 0"
				if (idx >= 0) {
					do {  // both words present and in the same state (the writer updates a, then b)
						wa = ldRelaxed64(&st[idx].a);
						wb = ldRelaxed64(&st[idx].b);
					} while ((wa >> 62) == 0 || (wa >> 62) != (wb >> 62));
				}
				const u32 done = __ballot_sync(0xFFFFFFFFu, (wa >> 62) == 2);
				const int first = __ffs(done) - 1;  // nearest tile that already knows its inclusive prefix
				u64 vA = wa & kScanValueMask, vB = wb & kScanValueMask;
				if (done != 0u && lane > first) {
					vA = 0;
					vB = 0;
				}
#pragma unroll
				for (int d = 16; d > 0; d >>= 1) {
					vA += __shfl_xor_sync(0xFFFFFFFFu, vA, d);
					vB += __shfl_xor_sync(0xFFFFFFFFu, vB, d);
				}
				runA += vA;
				runB += vB;
				if (done != 0u) break;
				look -= 32;
			}
			if (lane == 0) {
				stRelaxed64(&st[tile].a, (2ull << 62) | (runA + totA));
				stRelaxed64(&st[tile].b, (2ull << 62) | (runB + totB));
			}
		}
		if (lane == 0) {
			sPre[0] = runA;
			sPre[1] = runB;
		}
	}
	__syncthreads();
	preA = sPre[0];
	preB = sPre[1];
	__syncthreads();
}

}  // namespace cpvs
