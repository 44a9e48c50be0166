/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's shadow-DAG path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (cpvs_b200/csrc, include/cpvs_b200.h) never links, imports or falls back to it.
 *
 * What it is: the reference algorithm (depth map -> min/max pyramid -> SVO -> bottom-up subtree
 * merge -> pointer compression -> lookup, plus the container's top-level grid) written again from
 * the reference's behaviour, each function citing the reference file:line it follows. The one
 * deliberate change is in mergeLevel: the reference finds the first identical node by a nested
 * linear search (O(n*u)); this port finds the same first occurrence through a hash table keyed on
 * the full node tuple (O(n)), so it can be run at 16K^2 where the reference needs hours.
 *
 * Pinned: tests/test_oracle.py checks this port word-for-word against (a) the reference's own
 * known-answer tests (test/CompressedShadowUtilTest.cpp, test/CompressedShadowTest.cpp,
 * test/MinMaxTest.cpp) through the fixtures in tests/golden/, and (b) outputs of the unmodified
 * reference compiled into oracle/_ref/ (when present) on every synthetic generator. Parity is
 * therefore NOT "unpinned".
 *
 * Defined behaviour where the reference has none:
 *   N1  zTileNum is an argument, not a file-static.
 *   N2  levels are tracked by explicit node counts, so an SVO that stops early (a level whose
 *       nodes have no PARTIAL child) yields a valid, shorter DAG instead of reading out of bounds.
 *   N3  the container lookup uses the grid sentinels the C++ side writes (0x0FFFFFFF/0x0FFFFFFE).
 *   N5  lookup paths are clamped to [0, RES-1].
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

typedef uint32_t uint;

namespace {

enum NodeVisibility { SHADOW = 0, VISIBLE = 1, PARTIAL = 2 }; /* src/CompressedShadow.h:22-26 */
const uint NODE_SIZE = 9;                                     /* src/CompressedShadowUtil.h:9 */
const uint LEAF_SIZE = 17;                                    /* src/CompressedShadowUtil.h:10 */
const uint GRID_CELL_SHADOWED = 0xFFFFFFF;                    /* src/CompressedShadowContainer.cpp:8 */
const uint GRID_CELL_VISIBLE = 0xFFFFFFE;                     /* src/CompressedShadowContainer.cpp:9 */

/* ---- MinMaxHierarchy (src/MinMaxHierarchy.h:23-72, src/MinMaxHierarchy.cpp:9-97) ---------------- */
struct MinMax {
	int n;                                  /* side of level 0 */
	std::vector<float> root;                /* level 0: n*n depths (m_root, .cpp:10) */
	std::vector<std::vector<float> > levels; /* level k>=1 at [k-1]: (n>>k)^2 interleaved (min,max) */

	int numLevels() const { return (int)levels.size() + 1; } /* .h:60-62 */
	int side(int level) const { return n >> level; }
	float getMin(int level, size_t x, size_t y) const { /* .h:31-40 */
		if (level == 0) return root[y * n + x];
		return levels[level - 1][(y * side(level) + x) * 2 + 0];
	}
	float getMax(int level, size_t x, size_t y) const { /* .h:46-55 */
		if (level == 0) return root[y * n + x];
		return levels[level - 1][(y * side(level) + x) * 2 + 1];
	}
};

/* std::min / std::max semantics, reduction order pred(pred(a,b),pred(c,d)) (.cpp:29-33,46-47). */
inline float minOf(float a, float b) { return (b < a) ? b : a; }
inline float maxOf(float a, float b) { return (a < b) ? b : a; }

MinMax* buildMinMax(const float* depth, int n) {
	MinMax* mm = new MinMax;
	mm->n = n;
	mm->root.assign(depth, depth + (size_t)n * n);
	int numLevels = 0; /* ceil(log2(n)) for a power of two (.cpp:17) */
	while ((1 << numLevels) < n) ++numLevels;
	for (int k = 1; k <= numLevels; ++k) {
		const int s = n >> k, ps = s * 2;
		std::vector<float> lvl((size_t)s * s * 2);
		for (int y = 0; y < s; ++y)
			for (int x = 0; x < s; ++x) {
				float a0, b0, c0, d0, a1, b1, c1, d1;
				if (k == 1) { /* constructLevelFromRoot: both channels read channel 0 (.cpp:75-97) */
					a0 = a1 = depth[(size_t)(2 * y) * ps + 2 * x];
					b0 = b1 = depth[(size_t)(2 * y) * ps + 2 * x + 1];
					c0 = c1 = depth[(size_t)(2 * y + 1) * ps + 2 * x];
					d0 = d1 = depth[(size_t)(2 * y + 1) * ps + 2 * x + 1];
				} else { /* constructLevel (.cpp:51-73) */
					const std::vector<float>& p = mm->levels[k - 2];
					a0 = p[((size_t)(2 * y) * ps + 2 * x) * 2], a1 = p[((size_t)(2 * y) * ps + 2 * x) * 2 + 1];
					b0 = p[((size_t)(2 * y) * ps + 2 * x + 1) * 2], b1 = p[((size_t)(2 * y) * ps + 2 * x + 1) * 2 + 1];
					c0 = p[((size_t)(2 * y + 1) * ps + 2 * x) * 2], c1 = p[((size_t)(2 * y + 1) * ps + 2 * x) * 2 + 1];
					d0 = p[((size_t)(2 * y + 1) * ps + 2 * x + 1) * 2], d1 = p[((size_t)(2 * y + 1) * ps + 2 * x + 1) * 2 + 1];
				}
				lvl[((size_t)y * s + x) * 2 + 0] = minOf(minOf(a0, b0), minOf(c0, d0));
				lvl[((size_t)y * s + x) * 2 + 1] = maxOf(maxOf(a1, b1), maxOf(c1, d1));
			}
		mm->levels.push_back(std::move(lvl));
	}
	return mm;
}

/* ---- classification (src/CompressedShadowUtil.h:35-57, src/CompressedShadowUtil.cpp:16-99) ------- */
inline uint visible(float minZ, float maxZ, float minDepth, float maxDepth) { /* Util.h:35-44 */
	if (maxZ <= minDepth) return VISIBLE;
	if (minZ >= maxDepth) return SHADOW;
	return PARTIAL;
}
inline uint absoluteVisible(float minZ, float maxZ, float depth) { /* Util.h:51-57 */
	const float midZ = (minZ + maxZ) * 0.5f;
	return (midZ <= depth) ? VISIBLE : SHADOW;
}
inline uint levelHeight(const MinMax& mm, uint level, uint zTileNum) { /* Util.cpp:16-18 */
	return (uint)mm.side(level) * zTileNum;
}

uint createChildmask(const MinMax& mm, uint level, int ox, int oy, int oz, uint zTileNum) { /* Util.cpp:20-54 */
	const uint h = levelHeight(mm, level, zTileNum);
	uint childmask = 0;
	for (uint z = 0; z < 2; ++z)
		for (uint y = 0; y < 2; ++y)
			for (uint x = 0; x < 2; ++x) {
				const uint offZ = z + oz, offY = y + oy, offX = x + ox;
				const float mn = mm.getMin(level, offX, offY), mx = mm.getMax(level, offX, offY);
				uint bits;
				if (level > 0)
					bits = visible((float)offZ, (float)(offZ + 1), mn * h, mx * h);
				else
					bits = absoluteVisible((float)offZ, (float)(offZ + 1), mn * h);
				childmask |= bits << ((x | (y << 1) | (z << 2)) * 2);
			}
	return childmask;
}

uint64_t createLeafmask(const MinMax& mm, int ox, int oy, int oz, uint zTileNum) { /* Util.cpp:59-78 */
	const uint h = levelHeight(mm, 0, zTileNum);
	uint64_t leafmask = 0;
	uint index = 0;
	for (uint y = 0; y < 8; ++y)
		for (uint x = 0; x < 8; ++x, ++index) {
			const float d = mm.getMin(0, ox + x, oy + y);
			const uint64_t bit = absoluteVisible((float)oz, (float)(oz + 1), d * h);
			leafmask |= bit << index;
		}
	return leafmask;
}

/* 1x1x8 stack of 8x8x1 leafmasks for the level-2 node at (ox,oy,oz) (Util.cpp:80-99). */
uint createChildmask1x1x8(const MinMax& mm, int ox, int oy, int oz, uint zTileNum, uint64_t masks[8], uint* numMasks) {
	int cx = ox * 4, cy = oy * 4, cz = oz * 4;
	uint childmask = 0, k = 0;
	for (uint z = 0; z < 8; ++z, ++cz) {
		const uint64_t m = createLeafmask(mm, cx, cy, cz, zTileNum);
		if (m == 0xFFFFFFFFFFFFFFFFull)
			childmask |= 1u << (z * 2);
		else if (m != 0) {
			childmask |= 2u << (z * 2);
			masks[k++] = m;
		}
	}
	*numMasks = k;
	return childmask;
}

inline uint numChildren(uint mask) { return (uint)__builtin_popcount(mask & 0xAAAA); } /* Util.h:80-85 */

/* ---- CompressedShadow (src/CompressedShadow.cpp) --------------------------------------------- */
struct Coord {
	int x, y, z;
};

struct Shadow {
	uint numLevels;
	bool leafmasks;
	std::vector<uint> dag;
	std::vector<uint64_t> svoNodes;  /* per level: SVO nodes before merge */
	std::vector<uint64_t> uniqNodes; /* per level: nodes left after merge */

	bool useLeafmasks() const { return leafmasks && (numLevels - 3) >= 2; } /* .cpp:20-27 */
	uint minLevel() const { return useLeafmasks() ? 2 : 0; }                /* .cpp:30-32 */
	uint nodeSize(int level) const { return (useLeafmasks() && level == 2) ? LEAF_SIZE : NODE_SIZE; } /* .cpp:35-41 */
};

/* Uncompressed SVO, breadth first, top level first (src/CompressedShadow.cpp:87-190).
 * levelStart[l] / levelCount[l]: word offset and node count of level l (root = numLevels-2). */
void constructSvo(Shadow& s, const MinMax& mm, uint zTileIndex, uint zTileNum, std::vector<uint64_t>& levelStart,
		std::vector<uint64_t>& levelCount) {
	const int top = (int)s.numLevels - 2;
	levelStart.assign(s.numLevels - 1, 0);
	levelCount.assign(s.numLevels - 1, 0);
	std::vector<uint>& dag = s.dag;
	dag.clear();

	std::vector<Coord> coords(1, Coord{0, 0, (int)(zTileIndex * 2)});
	const int lastInner = s.useLeafmasks() ? 3 : 0;
	uint64_t levelOffset = 0;
	dag.resize(NODE_SIZE, 0);
	int level = top;
	/* The reference unrolls the root (.cpp:88-99); the loop body below is the same computation. */
	for (; level >= lastInner && !coords.empty(); --level) {
		const uint64_t n = coords.size();
		levelStart[level] = levelOffset;
		levelCount[level] = n;
		const uint childSize = (level > 0) ? s.nodeSize(level - 1) : 0;
		const uint64_t nextOffset = levelOffset + n * NODE_SIZE;
		uint64_t newChildren = 0;
		for (uint64_t i = 0; i < n; ++i) { /* pass 1: masks + count (.cpp:125-132) */
			const uint mask = createChildmask(mm, level, coords[i].x, coords[i].y, coords[i].z, zTileNum);
			dag[levelOffset + i * NODE_SIZE] = mask;
			newChildren += numChildren(mask);
		}
		if (level != 0) dag.resize(dag.size() + newChildren * childSize, 0); /* .cpp:134-135 */
		std::vector<Coord> next;
		next.reserve(newChildren);
		uint64_t progress = 0;
		for (uint64_t i = 0; i < n; ++i) { /* pass 2: pointers + child coordinates (.cpp:137-151) */
			const uint64_t nodeOffset = levelOffset + i * NODE_SIZE;
			const uint mask = dag[nodeOffset];
			uint k = 0;
			for (uint c = 0; c < 8; ++c) { /* getChildCoordinates (Util.cpp:101-117) */
				if (!(mask & (2u << (c * 2)))) continue;
				next.push_back(Coord{(coords[i].x + (int)(c & 1)) * 2, (coords[i].y + (int)((c >> 1) & 1)) * 2,
									 (coords[i].z + (int)((c >> 2) & 1)) * 2});
				dag[nodeOffset + 1 + k] = (uint)(nextOffset + progress + (uint64_t)k * childSize); /* .cpp:78-85 */
				++k;
			}
			progress += (uint64_t)k * childSize;
		}
		levelOffset = nextOffset;
		coords.swap(next);
	}
	if (s.useLeafmasks() && level == 2 && !coords.empty()) { /* constructLastLevels (.cpp:171-190) */
		levelStart[2] = levelOffset;
		levelCount[2] = coords.size();
		for (uint64_t i = 0; i < coords.size(); ++i) {
			uint64_t masks[8];
			uint k;
			const uint64_t nodeOffset = levelOffset + i * LEAF_SIZE;
			dag[nodeOffset] = createChildmask1x1x8(mm, coords[i].x, coords[i].y, coords[i].z, zTileNum, masks, &k);
			for (uint c = 0; c < k; ++c) {
				dag[nodeOffset + 1 + 2 * c] = (uint)masks[c];
				dag[nodeOffset + 2 + 2 * c] = (uint)(masks[c] >> 32);
			}
		}
	}
	s.svoNodes = levelCount;
}

/* cs::mergeLevel (src/CompressedShadowUtil.h:154-182): keep the first occurrence of every distinct
 * nodeSize-word tuple, in order; mapping[i] = word offset (inside the level) node i now lives at.
 * Restated with an open-addressing table over node indices; equality is the full tuple compare of
 * cs::isEqualSubtree (Util.h:137-145), so the result is the reference's, not a hash's. */
uint64_t mergeLevel(uint* level, uint64_t n, uint nodeSize, std::vector<uint>& mapping) {
	mapping.resize(n);
	uint64_t cap = 16;
	while (cap < n * 2) cap <<= 1;
	std::vector<uint> table(cap, 0xFFFFFFFFu); /* holds NEW node indices */
	uint64_t kept = 0;
	for (uint64_t i = 0; i < n; ++i) {
		const uint* node = level + i * nodeSize;
		uint64_t h = 0x9E3779B97F4A7C15ull;
		for (uint w = 0; w < nodeSize; ++w) {
			h ^= node[w];
			h *= 0xFF51AFD7ED558CCDull;
			h ^= h >> 29;
		}
		uint64_t slot = h & (cap - 1);
		for (;; slot = (slot + 1) & (cap - 1)) {
			const uint j = table[slot];
			if (j == 0xFFFFFFFFu) { /* first occurrence: append in place (kept <= i) */
				table[slot] = (uint)kept;
				if (kept != i) std::memmove(level + kept * nodeSize, node, nodeSize * sizeof(uint));
				mapping[i] = (uint)(kept * nodeSize);
				++kept;
				break;
			}
			if (std::memcmp(level + (uint64_t)j * nodeSize, node, nodeSize * sizeof(uint)) == 0) {
				mapping[i] = j * nodeSize;
				break;
			}
		}
	}
	/* the reference copies a zero-filled temp level back (.cpp:226,235): tail becomes zero */
	std::fill(level + kept * nodeSize, level + n * nodeSize, 0u);
	return kept;
}

/* mergeCommonSubtrees + updateParentPointers + removeUnusedNodes (src/CompressedShadow.cpp:215-304). */
void mergeCommonSubtrees(Shadow& s, const std::vector<uint64_t>& levelStart, const std::vector<uint64_t>& levelCount) {
	const int top = (int)s.numLevels - 2;
	std::vector<uint64_t> kept(levelCount);
	std::vector<uint> mapping;
	for (int level = (int)s.minLevel(); level < top; ++level) { /* bottom-up, root never merged (.cpp:221) */
		if (levelCount[level] == 0) continue;
		const uint ns = s.nodeSize(level);
		kept[level] = mergeLevel(&s.dag[levelStart[level]], levelCount[level], ns, mapping);
		/* updateParentPointers(level + 1) (.cpp:243-259) */
		const uint64_t childBase = levelStart[level];
		uint* parent = &s.dag[levelStart[level + 1]];
		for (uint64_t i = 0; i < levelCount[level + 1]; ++i)
			for (uint c = 1; c < NODE_SIZE; ++c) {
				uint& p = parent[i * NODE_SIZE + c];
				if (p != 0) p = (uint)(childBase + mapping[(p - childBase) / ns]);
			}
	}
	s.uniqNodes = kept;
	/* removeUnusedNodes (.cpp:271-304): keep the first kept[l] nodes of every level, shift pointers
	 * down by the words removed in front of their target. */
	std::vector<uint> out(s.dag.begin(), s.dag.begin() + NODE_SIZE);
	uint64_t correction = 0;
	for (int level = top - 1; level >= (int)s.minLevel(); --level) {
		if (levelCount[level] == 0) continue;
		const uint ns = s.nodeSize(level);
		const uint64_t size = levelCount[level] * ns, merged = kept[level] * ns;
		correction += size - merged;
		const uint* src = &s.dag[levelStart[level]];
		const bool isLeafLevel = s.useLeafmasks() && level == 2;
		for (uint64_t i = 0; i < merged; ++i) {
			uint v = src[i];
			/* pointers (not masks, not nulls, not leaf payload) move by `correction` (.cpp:262-269,290-297) */
			if (!isLeafLevel && level != (int)s.minLevel() && (i % NODE_SIZE != 0) && v != 0) v -= (uint)correction;
			out.push_back(v);
		}
	}
	s.dag.swap(out);
}

/* compress (src/CompressedShadow.cpp:314-392): node = mask + one word per PARTIAL child (leaf: two
 * words per PARTIAL slice); pointers re-addressed to the compacted offsets, top-down. */
void compress(Shadow& s) {
	const int top = (int)s.numLevels - 2;
	const int minLevel = (int)s.minLevel();
	std::vector<uint> out;
	std::vector<uint64_t> oldStart(s.numLevels - 1, 0), newStart(s.numLevels - 1, 0);
	uint64_t oldOffset = 0;
	std::vector<std::vector<uint> > newOffsetOf(s.numLevels - 1); /* per level: node index -> new word offset */
	for (int level = top; level >= minLevel; --level) {
		const uint64_t n = s.uniqNodes[level];
		if (n == 0) break;
		const uint ns = s.nodeSize(level);
		const bool leaf = s.useLeafmasks() && level == 2;
		oldStart[level] = oldOffset;
		newStart[level] = out.size();
		newOffsetOf[level].resize(n);
		for (uint64_t i = 0; i < n; ++i) {
			const uint* node = &s.dag[oldOffset + i * ns];
			const uint k = numChildren(node[0]);
			newOffsetOf[level][i] = (uint)out.size();
			out.insert(out.end(), node, node + 1 + (leaf ? 2 * k : k)); /* copyNodeInNewDag (.cpp:314-324) */
		}
		if (level != top) { /* patch the parent level's pointers (.cpp:362-380) */
			const int pl = level + 1;
			uint64_t pos = newStart[pl];
			for (uint64_t i = 0; i < s.uniqNodes[pl]; ++i) {
				const uint k = numChildren(out[pos]);
				for (uint c = 0; c < k; ++c) {
					const uint old = out[pos + 1 + c];
					if (old != 0) out[pos + 1 + c] = newOffsetOf[level][(old - oldOffset) / ns];
				}
				pos += 1 + k;
			}
		}
		oldOffset += n * ns;
	}
	s.dag.swap(out);
	s.dag.shrink_to_fit();
}

Shadow* createShadow(const MinMax& mm, uint zTileIndex, uint zTileNum, bool leafmasks) { /* .cpp:49-59 */
	Shadow* s = new Shadow;
	s->numLevels = (uint)mm.numLevels();
	s->leafmasks = leafmasks;
	std::vector<uint64_t> levelStart, levelCount;
	constructSvo(*s, mm, zTileIndex, zTileNum, levelStart, levelCount);
	mergeCommonSubtrees(*s, levelStart, levelCount);
	compress(*s);
	return s;
}

inline int totalVisibility(const uint* dag) { /* .cpp:66-72 */
	if (dag[0] == 0x5555) return VISIBLE;
	if (dag[0] == 0) return SHADOW;
	return PARTIAL;
}

inline uint childOffset(uint mask, uint childBits) { /* getChildOffset (.cpp:394-402), childBits = 2*index */
	return (uint)__builtin_popcount(mask & (0xAAAAu >> (16 - childBits)));
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cs::getPathFromNDC (Util.h:70-75) with the N5 clamp; resLevels = numLevels (+ grid levels). */
inline void pathFromNdc(float x, float y, float z, uint resLevels, int path[3]) {
	const int resolution = (1 << (resLevels - 1)) - 1;
	const float v[3] = {x, y, z};
	for (int i = 0; i < 3; ++i) {
		float f = v[i] + 1.0f;
		f *= 0.5f;
		f = f * (float)resolution;
		int p = (f != f) ? 0 : (f <= 0.f ? 0 : (f >= (float)resolution ? resolution : (int)f));
		path[i] = clampi(p, 0, resolution);
	}
}

/* Descent shared by CompressedShadow::traverse (.cpp:404-463) and traverse.cs:75-133. `dag` points
 * at the cell's first word; pointers are relative to it (the shader adds dagOffset, traverse.cs:107). */
int descend(const uint* dag, uint numLevels, bool leaf, const int path[3]) {
	uint64_t offset = 0;
	int level = (int)numLevels - 2;
	const int minLevel = leaf ? 3 : 0;
	for (; level >= minLevel; --level) {
		const int bit = 1 << level;
		const uint idx = ((path[0] & bit) ? 1 : 0) + ((path[1] & bit) ? 2 : 0) + ((path[2] & bit) ? 4 : 0);
		const uint mask = dag[offset];
		const uint vis = (mask >> (idx * 2)) & 3;
		if (vis == 1) return VISIBLE;
		if (vis == 0) return SHADOW;
		offset = dag[offset + 1 + childOffset(mask, idx * 2)];
	}
	if (leaf) {
		const uint idx = path[2] & 7;
		const uint mask = dag[offset];
		const uint vis = (mask >> (idx * 2)) & 3;
		if (vis == 1) return VISIBLE;
		if (vis == 0) return SHADOW;
		const uint64_t index = offset + childOffset(mask, idx * 2) * 2 + 1;
		const int mi = (path[0] & 7) + 8 * (path[1] & 7);
		const uint word = (mi < 32) ? dag[index] : dag[index + 1];
		return ((word >> (mi & 31)) & 1) ? VISIBLE : SHADOW;
	}
	return PARTIAL;
}

/* ---- CompressedShadowContainer (src/CompressedShadowContainer.{h,cpp}) + shader/traverse.cs ----- */
struct Container {
	uint length;
	std::vector<std::shared_ptr<Shadow> > cells; /* index z*len^2 + y*len + x (.h:34-46) */
	std::vector<uint> dag, grid;
	uint dagLevels, gridLevels;
};

void finalizeContainer(Container& c) {
	c.dag.clear();
	c.grid.clear();
	uint offset = 0;
	for (size_t i = 0; i < c.cells.size(); ++i) { /* combineDAGs (.cpp:52-69) + createTopLevelGrid (.cpp:71-91) */
		const Shadow& s = *c.cells[i];
		const int vis = totalVisibility(s.dag.data());
		c.grid.push_back(vis == SHADOW ? GRID_CELL_SHADOWED : (vis == VISIBLE ? GRID_CELL_VISIBLE : offset));
		offset += (uint)s.dag.size();
		c.dag.insert(c.dag.end(), s.dag.begin(), s.dag.end());
	}
	c.dagLevels = c.cells[0]->numLevels; /* .cpp:40 */
	c.gridLevels = 0;                    /* log8(#cells) (.cpp:42-43), exact for len = 2^k */
	while ((1u << c.gridLevels) < c.length) ++c.gridLevels;
}

int containerLookup(const Container& c, float x, float y, float z) { /* traverse.cs:41-48,75-133 */
	int path[3];
	pathFromNdc(x, y, z, c.dagLevels + c.gridLevels, path);
	const uint gridRes = 1u << c.gridLevels;
	const uint gx = path[0] >> (c.dagLevels - 1), gy = path[1] >> (c.dagLevels - 1), gz = path[2] >> (c.dagLevels - 1);
	const uint cell = c.grid[gz * gridRes * gridRes + gy * gridRes + gx];
	if (cell == GRID_CELL_SHADOWED) return SHADOW; /* N3: the values the C++ side writes */
	if (cell == GRID_CELL_VISIBLE) return VISIBLE;
	const bool leaf = c.cells[0]->useLeafmasks();
	return descend(&c.dag[cell], c.dagLevels, leaf, path);
}

}  // namespace

extern "C" {

void* orc_minmax_create(const float* depth, int n) { return buildMinMax(depth, n); }
void orc_minmax_destroy(void* h) { delete static_cast<MinMax*>(h); }
int orc_minmax_num_levels(void* h) { return static_cast<MinMax*>(h)->numLevels(); }
long orc_minmax_level(void* h, int level, float* out) {
	const MinMax* mm = static_cast<MinMax*>(h);
	const std::vector<float>& v = (level == 0) ? mm->root : mm->levels[level - 1];
	if (out) std::memcpy(out, v.data(), v.size() * sizeof(float));
	return (long)v.size();
}
unsigned orc_create_childmask(void* h, unsigned level, int x, int y, int z, unsigned zTileNum) {
	return createChildmask(*static_cast<MinMax*>(h), level, x, y, z, zTileNum);
}

void* orc_shadow_create(void* mm, unsigned zTileIndex, unsigned zTileNum, int leafmasks) {
	return createShadow(*static_cast<MinMax*>(mm), zTileIndex, zTileNum, leafmasks != 0);
}
void orc_shadow_destroy(void* h) { delete static_cast<Shadow*>(h); }
unsigned orc_shadow_num_levels(void* h) { return static_cast<Shadow*>(h)->numLevels; }
long orc_shadow_words(void* h) { return (long)static_cast<Shadow*>(h)->dag.size(); }
int orc_shadow_total_visibility(void* h) { return totalVisibility(static_cast<Shadow*>(h)->dag.data()); }
void orc_shadow_copy_dag(void* h, uint32_t* out) {
	const Shadow* s = static_cast<Shadow*>(h);
	std::memcpy(out, s->dag.data(), s->dag.size() * sizeof(uint32_t));
}
/* svo / uniq: numLevels-1 entries each, index = level. */
void orc_shadow_level_counts(void* h, uint64_t* svo, uint64_t* uniq) {
	const Shadow* s = static_cast<Shadow*>(h);
	for (uint l = 0; l + 1 < s->numLevels; ++l) {
		svo[l] = s->svoNodes[l];
		uniq[l] = s->uniqNodes[l];
	}
}
void orc_shadow_traverse(void* h, const float* ndc, long count, int tryLeafmasks, uint8_t* out) {
	const Shadow* s = static_cast<Shadow*>(h);
	for (long i = 0; i < count; ++i) {
		int path[3];
		pathFromNdc(ndc[3 * i], ndc[3 * i + 1], ndc[3 * i + 2], s->numLevels, path);
		out[i] = (uint8_t)descend(s->dag.data(), s->numLevels, tryLeafmasks != 0, path);
	}
}

/* Uncompressed SVO in the reference's 9/17-word layout; levelOffsets: numLevels-1 entries. */
long orc_svo(void* mm, unsigned zTileIndex, unsigned zTileNum, int leafmasks, uint32_t* out, uint32_t* levelOffsets) {
	Shadow s;
	s.numLevels = (uint)static_cast<MinMax*>(mm)->numLevels();
	s.leafmasks = leafmasks != 0;
	std::vector<uint64_t> levelStart, levelCount;
	constructSvo(s, *static_cast<MinMax*>(mm), zTileIndex, zTileNum, levelStart, levelCount);
	if (levelOffsets) {
		for (uint l = 0; l + 1 < s.numLevels; ++l) levelOffsets[l] = (uint32_t)levelStart[l];
		/* the reference also records the end of the leaf level one slot below it (.cpp:164) */
		if (s.useLeafmasks() && levelCount[2] > 0) levelOffsets[1] = (uint32_t)s.dag.size();
	}
	if (out) std::memcpy(out, s.dag.data(), s.dag.size() * sizeof(uint32_t));
	return (long)s.dag.size();
}

unsigned orc_merge_level(const uint32_t* level, long words, unsigned nodeSize, uint32_t* merged, uint32_t* mapping) {
	std::vector<uint> tmp(level, level + words), map;
	const uint64_t kept = mergeLevel(tmp.data(), words / nodeSize, nodeSize, map);
	std::memcpy(merged, tmp.data(), words * sizeof(uint32_t));
	std::memcpy(mapping, map.data(), map.size() * sizeof(uint32_t));
	return (unsigned)kept;
}

void* orc_container_create(unsigned length) {
	Container* c = new Container;
	c->length = length;
	c->cells.resize((size_t)length * length * length);
	return c;
}
void orc_container_destroy(void* h) { delete static_cast<Container*>(h); }
/* The container shares ownership with the caller's handle (the handle stays valid). */
void orc_container_set(void* h, void* shadow, unsigned x, unsigned y, unsigned z) {
	Container* c = static_cast<Container*>(h);
	Shadow* copy = new Shadow(*static_cast<Shadow*>(shadow));
	c->cells[(size_t)z * c->length * c->length + (size_t)y * c->length + x].reset(copy);
}
void orc_container_finalize(void* h) { finalizeContainer(*static_cast<Container*>(h)); }
long orc_container_words(void* h) { return (long)static_cast<Container*>(h)->dag.size(); }
void orc_container_copy(void* h, uint32_t* dag, uint32_t* grid) {
	const Container* c = static_cast<Container*>(h);
	if (dag) std::memcpy(dag, c->dag.data(), c->dag.size() * sizeof(uint32_t));
	if (grid) std::memcpy(grid, c->grid.data(), c->grid.size() * sizeof(uint32_t));
}
void orc_container_lookup_ndc(void* h, const float* ndc, long count, uint8_t* out) {
	const Container* c = static_cast<Container*>(h);
	for (long i = 0; i < count; ++i) out[i] = (uint8_t)containerLookup(*c, ndc[3 * i], ndc[3 * i + 1], ndc[3 * i + 2]);
}
/* traverse.cs main() (:135-149): pos = rgba32f texels (xyz used), m = column-major mat4 as glm stores
 * it; product and divide in glm's order ((m0*x + m1*y) + (m2*z + m3*w), then /w). out: r8 texels, 0 or 255. */
void orc_container_evaluate(void* h, const float* pos, long width, long height, const float* m, uint8_t* out) {
	const Container* c = static_cast<Container*>(h);
	for (long i = 0; i < width * height; ++i) {
		const float x = pos[4 * i], y = pos[4 * i + 1], z = pos[4 * i + 2], w = 1.0f;
		float p[4];
		for (int r = 0; r < 4; ++r) p[r] = (m[0 + r] * x + m[4 + r] * y) + (m[8 + r] * z + m[12 + r] * w);
		out[i] = containerLookup(*c, p[0] / p[3], p[1] / p[3], p[2] / p[3]) == VISIBLE ? 255 : 0;
	}
}

} /* extern "C" */
