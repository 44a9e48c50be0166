/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * C-ABI bridge over the UNMODIFIED reference classes, compiled together with the reference's own
 * translation units taken in place from /root/reference (see oracle/Makefile). The result,
 * oracle/_ref/libcpvs_ref.so, is the ground truth the CPU restatement (oracle/oracle_port.cpp)
 * and the CUDA path are pinned against, and the `kind: "reference"` CPU baseline of bench.py.
 *
 * Wrapped reference entry points:
 *   MinMaxHierarchy::MinMaxHierarchy            src/MinMaxHierarchy.cpp:9-27
 *   MinMaxHierarchy::getLevel / getNumLevels    src/MinMaxHierarchy.h:60-72
 *   CompressedShadow::create                    src/CompressedShadow.cpp:49-59
 *   CompressedShadow::constructSvo (private)    src/CompressedShadow.cpp:87-169
 *   CompressedShadow::traverse                  src/CompressedShadow.cpp:404-463
 *   CompressedShadow::getTotalVisibility        src/CompressedShadow.cpp:66-72
 *   cs::createChildmask                         src/CompressedShadowUtil.cpp:20-54
 *   cs::mergeLevel                              src/CompressedShadowUtil.h:154-182
 */
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>

#define private public /* reach constructSvo / m_dag for phase-level parity */
#include "CompressedShadow.h"
#undef private
#include "CompressedShadowUtil.h"
#include "MinMaxHierarchy.h"
#include "Image.h"

namespace {
struct RefMinMax {
	MinMaxHierarchy mm;
	explicit RefMinMax(const ImageF& img) : mm(img) {}
};
struct RefShadow {
	std::unique_ptr<CompressedShadow> cs;
};
ImageF makeImage(const float* depth, int n) {
	ImageF img(n, n, 1);
	img.setAll(depth);
	return img;
}
}

extern "C" {

int ref_leafmasks_compiled() {
#ifdef CPVS_REF_NOLEAF
	return 0;
#else
	return 1;
#endif
}

void* ref_minmax_create(const float* depth, int n) {
	return new RefMinMax(makeImage(depth, n));
}
void ref_minmax_destroy(void* h) { delete static_cast<RefMinMax*>(h); }
int ref_minmax_num_levels(void* h) { return static_cast<RefMinMax*>(h)->mm.getNumLevels(); }

/* Copies level `level` (level 0: n*n floats; level k: (n>>k)^2 interleaved (min,max) pairs). */
long ref_minmax_level(void* h, int level, float* out) {
	const ImageF* img = static_cast<RefMinMax*>(h)->mm.getLevel(level);
	const size_t count = img->getWidth() * img->getHeight() * img->getNumChannels();
	if (out) std::memcpy(out, img->data(), count * sizeof(float));
	return static_cast<long>(count);
}

unsigned ref_create_childmask(void* h, unsigned level, int x, int y, int z, unsigned zTileNum) {
	cs::setDepthOffset(zTileNum);
	return cs::createChildmask(static_cast<RefMinMax*>(h)->mm, level, ivec3(x, y, z));
}

void* ref_shadow_create(void* mmHandle, unsigned zTileIndex, unsigned zTileNum) {
	RefShadow* s = new RefShadow;
	s->cs = CompressedShadow::create(static_cast<RefMinMax*>(mmHandle)->mm, zTileIndex, zTileNum);
	return s;
}
void ref_shadow_destroy(void* h) { delete static_cast<RefShadow*>(h); }
unsigned ref_shadow_num_levels(void* h) { return static_cast<RefShadow*>(h)->cs->getNumLevels(); }
long ref_shadow_words(void* h) { return static_cast<long>(static_cast<RefShadow*>(h)->cs->getDAG().size()); }
int ref_shadow_total_visibility(void* h) { return static_cast<RefShadow*>(h)->cs->getTotalVisibility(); }
void ref_shadow_copy_dag(void* h, uint32_t* out) {
	const vector<uint>& dag = static_cast<RefShadow*>(h)->cs->getDAG();
	std::memcpy(out, dag.data(), dag.size() * sizeof(uint32_t));
}

/* ndc: count*3 floats; out: count bytes holding NodeVisibility (0 shadow, 1 visible, 2 partial). */
void ref_shadow_traverse(void* h, const float* ndc, long count, int tryLeafmasks, uint8_t* out) {
	CompressedShadow* cs = static_cast<RefShadow*>(h)->cs.get();
	for (long i = 0; i < count; ++i)
		out[i] = static_cast<uint8_t>(
			cs->traverse(vec3(ndc[3 * i], ndc[3 * i + 1], ndc[3 * i + 2]), tryLeafmasks != 0));
}

/* Same, split over `threads` host threads (traverse is read-only on the DAG). */
void ref_shadow_traverse_mt(void* h, const float* ndc, long count, int tryLeafmasks, uint8_t* out,
		int threads) {
	if (threads <= 1) return ref_shadow_traverse(h, ndc, count, tryLeafmasks, out);
	vector<std::thread> pool;
	const long chunk = (count + threads - 1) / threads;
	for (int t = 0; t < threads; ++t) {
		const long b = t * chunk, e = std::min(count, b + chunk);
		if (b >= e) break;
		pool.emplace_back([=]() { ref_shadow_traverse(h, ndc + 3 * b, e - b, tryLeafmasks, out + b); });
	}
	for (auto& th : pool) th.join();
}

/* Uncompressed SVO straight out of constructSvo (before merge / compress).
 * levelOffsets must hold numLevels-1 entries. Returns the SVO word count; call with out == NULL
 * first to size the buffer. */
long ref_svo(void* mmHandle, unsigned zTileIndex, unsigned zTileNum, uint32_t* out, uint32_t* levelOffsets) {
	const MinMaxHierarchy& mm = static_cast<RefMinMax*>(mmHandle)->mm;
	CompressedShadow cs(mm.getNumLevels());
	cs::setDepthOffset(zTileNum);
	vector<uint> levels = cs.constructSvo(mm, ivec3(0, 0, zTileIndex * 2));
	if (levelOffsets) std::memcpy(levelOffsets, levels.data(), levels.size() * sizeof(uint32_t));
	if (out) std::memcpy(out, cs.m_dag.data(), cs.m_dag.size() * sizeof(uint32_t));
	return static_cast<long>(cs.m_dag.size());
}

/* cs::mergeLevel on a caller-supplied level; mapping[i] = new word offset of node i. */
unsigned ref_merge_level(const uint32_t* level, long words, unsigned nodeSize, uint32_t* merged, uint32_t* mapping) {
	vector<uint> in(level, level + words), out(words, 0);
	uint left = 0;
	auto map = cs::mergeLevel(in.begin(), in.end(), out.begin(), nodeSize, &left);
	std::memcpy(merged, out.data(), words * sizeof(uint32_t));
	for (long i = 0; i < words / nodeSize; ++i) mapping[i] = map[i * nodeSize];
	return left;
}

/* Wall-clock of the reference build on this host: pyramid (4 threads inside) + create (1 thread). */
void ref_time_build(const float* depth, int n, unsigned zTileIndex, unsigned zTileNum, double* msMinMax,
		double* msCreate, long* words) {
	using clk = std::chrono::steady_clock;
	ImageF img = makeImage(depth, n);
	auto t0 = clk::now();
	MinMaxHierarchy mm(img);
	auto t1 = clk::now();
	auto cs = CompressedShadow::create(mm, zTileIndex, zTileNum);
	auto t2 = clk::now();
	*msMinMax = std::chrono::duration<double, std::milli>(t1 - t0).count();
	*msCreate = std::chrono::duration<double, std::milli>(t2 - t1).count();
	*words = static_cast<long>(cs->getDAG().size());
}

} /* extern "C" */
