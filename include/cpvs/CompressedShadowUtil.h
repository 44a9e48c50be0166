// namespace cs helpers with the reference's signatures (src/CompressedShadowUtil.h:8-183).
// createChildmask runs the device classification for one node (cpvs_minmax_childmask); the rest are
// the small pure functions the reference keeps inline. mergeLevel is provided for source compatibility
// with callers of the header; the builder merges levels on the device (cpvs_b200/csrc/merge.cu).
#ifndef CPVS_FACADE_COMPRESSED_SHADOW_UTIL_H
#define CPVS_FACADE_COMPRESSED_SHADOW_UTIL_H

#include <iterator>
#include <map>

#include "CompressedShadow.h"
#include "MinMaxHierarchy.h"
#include "cpvs.h"

namespace cs {
constexpr uint NODE_SIZE = 9;
constexpr uint LEAF_SIZE = 17;

namespace detail {
inline uint& depthOffset() {
	static uint value = 1;
	return value;
}
}  // namespace detail

// The reference keeps zTileNum in a file-static; kept here only for createChildmask.
inline void setDepthOffset(uint off) { detail::depthOffset() = off; }

inline uint createChildmask(const MinMaxHierarchy& minMax, uint level, const ivec3& offset) {
	uint32_t mask = 0;
	cpvs_facade::check(cpvs_minmax_childmask(minMax.handle(), level, (uint)offset.x, (uint)offset.y, (uint)offset.z, detail::depthOffset(), &mask));
	return mask;
}

inline vector<ivec3> getChildCoordinates(uint childmask, const ivec3& parentOffset) {
	vector<ivec3> result;
	for (uint i = 0; i < 8; ++i)
		if (childmask & (2u << (2 * i)))
			result.emplace_back(ivec3((parentOffset.x + (int)(i & 1)) * 2, (parentOffset.y + (int)((i >> 1) & 1)) * 2, (parentOffset.z + (int)(i >> 2)) * 2));
	return result;
}

inline size_t getResolution(size_t numLevels) { return (size_t)1 << (numLevels - 1); }
inline uint getNumChildren(uint childmask) { return POPCOUNT(childmask & 0xAAAA); }
inline bool isPartial(uint childmask, uint childIndex) { return (childmask >> (childIndex * 2 + 1)) & 1u; }
inline bool hasPartialChildren(uint childmask) { return getNumChildren(childmask) != 0; }
inline bool isCompletelyVisible(uint childmask) { return childmask == 0x5555; }
inline bool isVisible(uint childmask, uint childIndex) { return (childmask >> (childIndex * 2)) & 1u; }
inline bool isCompletelyShadowed(uint childmask) { return childmask == 0x0; }
inline bool isShadowed(uint childmask, uint childIndex) { return !isVisible(childmask, childIndex) && !isPartial(childmask, childIndex); }

template <typename It1, typename It2>
inline bool isEqualSubtree(It1 leftNode, It2 rightNode, uint nodeSize) {
	for (uint i = 0; i < nodeSize; ++i, ++leftNode, ++rightNode)
		if (*leftNode != *rightNode) return false;
	return true;
}

// First occurrence of every distinct node kept in order; result[oldWordOffset] = newWordOffset.
template <typename ItOld, typename ItNew>
unordered_map<uint, uint> mergeLevel(ItOld oldBegin, ItOld oldEnd, ItNew newBegin, uint nodeSize, uint* numNodesLeft) {
	unordered_map<uint, uint> result;
	std::map<vector<uint>, uint> seen;
	uint kept = 0, i = 0;
	for (ItOld it = oldBegin; it != oldEnd; std::advance(it, nodeSize), i += nodeSize) {
		ItOld last = it;
		std::advance(last, nodeSize);
		vector<uint> node(it, last);
		auto found = seen.find(node);
		if (found == seen.end()) {
			found = seen.emplace(node, kept * nodeSize).first;
			ItNew dst = newBegin;
			std::advance(dst, kept * nodeSize);
			std::copy(node.begin(), node.end(), dst);
			++kept;
		}
		result[i] = found->second;
	}
	*numNodesLeft = kept;
	return result;
}
}  // namespace cs

#endif
