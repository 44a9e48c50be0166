// MinMaxHierarchy with the reference's interface (src/MinMaxHierarchy.h:23-72), built on the GPU by
// cpvs_minmax_build. Levels are fetched to the host on first access (getMin/getMax are test helpers
// in the reference; the builder itself reads the device copy).
#ifndef CPVS_FACADE_MIN_MAX_HIERARCHY_H
#define CPVS_FACADE_MIN_MAX_HIERARCHY_H

#include "Image.h"
#include "cpvs.h"

class MinMaxHierarchy {
public:
	// zTileNum: the slicing the hierarchy's builds will use (createShadowTiles: one hierarchy, numSlices creates) -- a speed hint
	explicit MinMaxHierarchy(const ImageF& orig, cpvs_ctx* ctx = nullptr, uint zTileNum = 1) : m_ctx(ctx ? ctx : cpvs_facade::defaultContext()) {
		if (orig.getWidth() != orig.getHeight() || orig.getNumChannels() != 1)
			throw cpvs_facade::Error(CPVS_EINVAL, "MinMaxHierarchy: the image must be square with one channel");
		cpvs_facade::check(cpvs_minmax_build_tiled(m_ctx, orig.data(), (int)orig.getWidth(), CPVS_MEM_HOST, zTileNum, &m_handle));
		m_levels.resize(cpvs_minmax_num_levels(m_handle));
	}
	~MinMaxHierarchy() { cpvs_minmax_destroy(m_handle); }
	MinMaxHierarchy(const MinMaxHierarchy&) = delete;
	MinMaxHierarchy& operator=(const MinMaxHierarchy&) = delete;
	MinMaxHierarchy(MinMaxHierarchy&& o) : m_ctx(o.m_ctx), m_handle(o.m_handle), m_levels(std::move(o.m_levels)) { o.m_handle = nullptr; }

	float getMin(size_t level, size_t x, size_t y) const { return getLevel(level)->get(x, y, 0); }
	float getMax(size_t level, size_t x, size_t y) const { return getLevel(level)->get(x, y, level == 0 ? 0 : 1); }
	int getNumLevels() const { return (int)m_levels.size(); }

	const ImageF* getLevel(size_t level) const {
		unique_ptr<ImageF>& img = m_levels.at(level);
		if (!img) {
			const size_t side = (size_t)cpvs_minmax_size(m_handle) >> level;
			img.reset(new ImageF(side, side, level == 0 ? 1 : 2));
			cpvs_facade::check(cpvs_minmax_level(m_handle, (int)level, img->data()));
		}
		return img.get();
	}

	cpvs_minmax* handle() const { return m_handle; }
	cpvs_ctx* context() const { return m_ctx; }

private:
	cpvs_ctx* m_ctx;
	cpvs_minmax* m_handle = nullptr;
	mutable vector<unique_ptr<ImageF>> m_levels;
};

#endif
