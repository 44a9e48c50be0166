// Host image, interleaved channels, row-major: the interface of the reference's src/Image.h:10-69.
#ifndef CPVS_FACADE_IMAGE_H
#define CPVS_FACADE_IMAGE_H

#include "cpvs.h"

template <typename T>
class Image {
public:
	Image(size_t width, size_t height, size_t numChannels)
		: m_width(width), m_height(height), m_numChannels((int)numChannels), m_values(width * height * numChannels) {}

	void setAll(const T* ptr) { m_values.assign(ptr, ptr + m_width * m_height * m_numChannels); }
	void setAll(const vector<T>& vec) { m_values.assign(vec.begin(), vec.end()); }
	void setAll(typename vector<T>::iterator begin, typename vector<T>::iterator end) { m_values.assign(begin, end); }

	T get(size_t x, size_t y, size_t channel) const { return m_values[(y * m_width + x) * m_numChannels + channel]; }
	void set(size_t x, size_t y, size_t channel, T val) { m_values[(y * m_width + x) * m_numChannels + channel] = val; }

	const T* data() const { return m_values.data(); }
	T* data() { return m_values.data(); }

	size_t getNumChannels() const { return m_numChannels; }
	size_t getWidth() const { return m_width; }
	size_t getHeight() const { return m_height; }

private:
	size_t m_width, m_height;
	int m_numChannels;
	vector<T> m_values;
};

using ImageF = Image<float>;

#endif
