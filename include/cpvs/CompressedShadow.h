// CompressedShadow with the reference's interface (src/CompressedShadow.h:20-126) over the C ABI.
#ifndef CPVS_FACADE_COMPRESSED_SHADOW_H
#define CPVS_FACADE_COMPRESSED_SHADOW_H

#include "MinMaxHierarchy.h"
#include "ShadowMap.h"
#include "cpvs.h"

class CompressedShadow {
public:
	enum NodeVisibility { SHADOW = 0, VISIBLE = 1, PARTIAL = 2 };

	~CompressedShadow() { cpvs_shadow_destroy(m_handle); }
	CompressedShadow(const CompressedShadow&) = delete;
	CompressedShadow& operator=(const CompressedShadow&) = delete;

	// src/CompressedShadow.h:48-49. `leafmasks` replaces the reference's compile-time LEAFMASKS switch.
	static unique_ptr<CompressedShadow> create(const MinMaxHierarchy& minMax, uint zTileIndex = 0, uint zTileNum = 1, bool leafmasks = true) {
		cpvs_shadow* h = nullptr;
		cpvs_facade::check(cpvs_shadow_create(minMax.context(), minMax.handle(), zTileIndex, zTileNum, leafmasks ? 1 : 0, &h));
		return unique_ptr<CompressedShadow>(new CompressedShadow(h));
	}

	// src/CompressedShadow.h:55-56 (src/CompressedShadow.cpp:61-64): the hierarchy is a temporary of the call.
	static unique_ptr<CompressedShadow> create(const ShadowMap* shadowMap, uint zTileIndex = 0, uint zTileNum = 1, bool leafmasks = true,
			cpvs_ctx* ctx = nullptr) {
		cpvs_shadow* h = nullptr;
		cpvs_facade::check(cpvs_shadow_create_from_depth(ctx ? ctx : cpvs_facade::defaultContext(), shadowMap->data(), (int)shadowMap->getSize(),
				shadowMap->memoryKind(), zTileIndex, zTileNum, leafmasks ? 1 : 0, &h));
		return unique_ptr<CompressedShadow>(new CompressedShadow(h));
	}

	// src/CompressedShadow.h:64 -- one point; use traverse(points, count, ...) for batches.
	NodeVisibility traverse(const vec3 position, bool tryLeafmasks = true) {
		const float p[3] = {position.x, position.y, position.z};
		uint8_t out = 0;
		cpvs_facade::check(cpvs_shadow_lookup_ndc(m_handle, p, 1, CPVS_MEM_HOST, tryLeafmasks ? 1 : 0, &out));
		return static_cast<NodeVisibility>(out);
	}
	void traverse(const float* ndcXyz, int64_t count, uint8_t* out, bool tryLeafmasks = true) {
		cpvs_facade::check(cpvs_shadow_lookup_ndc(m_handle, ndcXyz, count, CPVS_MEM_HOST, tryLeafmasks ? 1 : 0, out));
	}

	NodeVisibility getTotalVisibility() const { return static_cast<NodeVisibility>(m_info.total_visibility); }
	uint getNumLevels() const { return m_info.num_levels; }

	// src/CompressedShadow.h:75 -- host copy, fetched on first use.
	const vector<uint>& getDAG() const {
		if (m_dag.empty() && m_info.words) {
			m_dag.resize(m_info.words);
			cpvs_facade::check(cpvs_shadow_copy_dag(m_handle, m_dag.data()));
		}
		return m_dag;
	}

	const cpvs_shadow_info& info() const { return m_info; }
	cpvs_shadow* handle() const { return m_handle; }

private:
	explicit CompressedShadow(cpvs_shadow* h) : m_handle(h) { cpvs_facade::check(cpvs_shadow_info_get(h, &m_info)); }

	cpvs_shadow* m_handle;
	cpvs_shadow_info m_info;
	mutable vector<uint> m_dag;
};

#endif
