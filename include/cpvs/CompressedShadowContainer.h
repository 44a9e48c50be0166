// CompressedShadowContainer with the reference's interface (src/CompressedShadowContainer.h:18-92).
// evaluate() takes host or device arrays (positions = width*height rgba32f texels, visibilities = r8 texels) or CUDA surface
// objects over the textures' arrays; with -DCPVS_WITH_GL also the GL texture names themselves (CUDA-GL interop).
#ifndef CPVS_FACADE_COMPRESSED_SHADOW_CONTAINER_H
#define CPVS_FACADE_COMPRESSED_SHADOW_CONTAINER_H

#include "CompressedShadow.h"
#include "cpvs.h"

class CompressedShadowContainer {
public:
	explicit CompressedShadowContainer(uint length, cpvs_ctx* ctx = nullptr) : m_length(length), m_data(length * length * length) {
		cpvs_facade::check(cpvs_container_create(ctx ? ctx : cpvs_facade::defaultContext(), length, &m_handle));
	}
	explicit CompressedShadowContainer(unique_ptr<CompressedShadow> shadow, cpvs_ctx* ctx = nullptr) : CompressedShadowContainer(1u, ctx) {
		set(std::move(shadow), 0, 0, 0);
	}
	~CompressedShadowContainer() { cpvs_container_destroy(m_handle); }
	CompressedShadowContainer(const CompressedShadowContainer&) = delete;
	CompressedShadowContainer& operator=(const CompressedShadowContainer&) = delete;

	void set(unique_ptr<CompressedShadow> shadow, uint x, uint y, uint z) {
		cpvs_facade::check(cpvs_container_set(m_handle, shadow->handle(), x, y, z));
		m_data.at(((size_t)z * m_length + y) * m_length + x) = std::move(shadow);
	}
	const CompressedShadow* get(uint x, uint y, uint z) const { return m_data.at(((size_t)z * m_length + y) * m_length + x).get(); }

	void copyToGPU() { cpvs_facade::check(cpvs_container_finalize(m_handle)); }
	void freeOnCPU() { vector<unique_ptr<CompressedShadow>>().swap(m_data); }
	void moveToGPU() {
		copyToGPU();
		freeOnCPU();
	}
	void setFilterSize(uint size) { cpvs_facade::check(cpvs_container_set_filter_size(m_handle, size)); }

	void evaluate(const float* positionsWS, uint width, uint height, const mat4& lightViewProj, uint8_t* visibilities, int mem = CPVS_MEM_HOST) {
		cpvs_facade::check(cpvs_container_evaluate(m_handle, positionsWS, width, height, mem, cpvs_facade::matrixData(lightViewProj), visibilities));
	}
	// the rgba32f position texture and the r8 visibility texture as cudaSurfaceObject_t (src/CompressedShadowContainer.cpp:100-107)
	void evaluate(unsigned long long positionsSurface, unsigned long long visibilitiesSurface, uint width, uint height, const mat4& lightViewProj) {
		cpvs_facade::check(cpvs_container_evaluate_surface(m_handle, positionsSurface, visibilitiesSurface, width, height,
				cpvs_facade::matrixData(lightViewProj)));
	}
#ifdef CPVS_WITH_GL
	// src/CompressedShadowContainer.h:59-60 with GL texture names (the reference passes Texture2D objects)
	void evaluate(unsigned int positionsTexture, unsigned int visibilitiesTexture, uint width, uint height, const mat4& lightViewProj) {
		cpvs_facade::check(cpvs_container_evaluate_gl(m_handle, positionsTexture, visibilitiesTexture, width, height,
				cpvs_facade::matrixData(lightViewProj)));
	}
#endif
	void lookupNdc(const float* ndcXyz, int64_t count, uint8_t* out, int mem = CPVS_MEM_HOST) {
		cpvs_facade::check(cpvs_container_lookup_ndc(m_handle, ndcXyz, count, mem, out));
	}

	cpvs_container* handle() const { return m_handle; }

private:
	uint m_length;
	vector<unique_ptr<CompressedShadow>> m_data;
	cpvs_container* m_handle = nullptr;
};

#endif
