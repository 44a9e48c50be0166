// ShadowMap with the reference's interface (src/ShadowMap.h:9-33) for CompressedShadow::create(const ShadowMap*, ...)
// (src/CompressedShadow.h:55-56, src/CompressedShadow.cpp:61-64). The reference wraps the GL depth texture the light pass
// rendered and reads it back with glGetTexImage (src/ShadowMap.cpp:23-30); here the depth map is a host image or -- the
// case a CUDA renderer or the CUDA-GL interop produces -- device memory, which the build then never leaves.
#ifndef CPVS_FACADE_SHADOW_MAP_H
#define CPVS_FACADE_SHADOW_MAP_H

#include "Image.h"
#include "cpvs.h"

class ShadowMap {
public:
	// a depth image in host memory (what createImageF() of the reference returns)
	explicit ShadowMap(shared_ptr<ImageF> depth) : m_host(std::move(depth)), m_device(nullptr), m_size(m_host ? m_host->getWidth() : 0) {
		if (!m_host || m_host->getWidth() != m_host->getHeight() || m_host->getNumChannels() != 1 || !isPowerOfTwo((int)m_size))
			throw cpvs_facade::Error(CPVS_EINVAL, "ShadowMap: the depth image must be square, one channel, side a power of two");
	}
	// size x size float32 depths in device memory (borrowed: must outlive the builds made from it)
	ShadowMap(const float* deviceDepth, size_t size) : m_device(deviceDepth), m_size(size) {
		if (!deviceDepth || !isPowerOfTwo((int)size)) throw cpvs_facade::Error(CPVS_EINVAL, "ShadowMap: side must be a power of two");
	}

	// src/ShadowMap.h:26 -- only for host-side maps (a device map is consumed in place)
	ImageF createImageF() const {
		if (!m_host) throw cpvs_facade::Error(CPVS_EINVAL, "ShadowMap::createImageF: the depth map lives in device memory");
		return *m_host;
	}

	size_t getSize() const { return m_size; }
	const float* data() const { return m_host ? m_host->data() : m_device; }
	int memoryKind() const { return m_host ? CPVS_MEM_HOST : CPVS_MEM_DEVICE; }

private:
	shared_ptr<ImageF> m_host;
	const float* m_device;
	size_t m_size;
};

#endif
