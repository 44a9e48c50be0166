// C++ facade over the cpvs_b200 C ABI: common types.
//
// Stands in for the reference's src/cpvs.h (type aliases only -- no GL). With glm on the include path
// (the reference vendors glm 0.9.6) vec3/ivec3/mat4 are glm's, so reference call sites compile
// unchanged; without it minimal value types with the same member names are provided.
#ifndef CPVS_FACADE_CPVS_H
#define CPVS_FACADE_CPVS_H

#include <cassert>
#include <cmath>
#include <cstdint>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../cpvs_b200.h"

using std::shared_ptr;
using std::string;
using std::unique_ptr;
using std::unordered_map;
using std::vector;

using uint = unsigned int;
using uint64 = uint64_t;

#if defined(__has_include)
#if __has_include(<glm/glm.hpp>)
#define CPVS_FACADE_HAVE_GLM 1
#endif
#endif

#ifdef CPVS_FACADE_HAVE_GLM
#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>
using mat4 = glm::mat4;
using vec2 = glm::vec2;
using vec3 = glm::vec3;
using vec4 = glm::vec4;
using ivec2 = glm::ivec2;
using ivec3 = glm::ivec3;
#else
struct vec3 {
	float x, y, z;
	vec3() : x(0), y(0), z(0) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct ivec3 {
	int x, y, z;
	ivec3() : x(0), y(0), z(0) {}
	ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
	bool operator==(const ivec3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct mat4 {
	float m[16];  // column-major, as glm::value_ptr
};
#endif

#define POPCOUNT(x) __builtin_popcount(x)

inline constexpr bool isPowerOfTwo(int x) { return !(x & (x - 1)); }

namespace cpvs_facade {

// Thrown where the reference asserts or terminates ("Errors").
class Error : public std::runtime_error {
public:
	Error(int code, const char* what) : std::runtime_error(what), m_code(code) {}
	int code() const { return m_code; }

private:
	int m_code;
};

inline void check(int rc) {
	if (rc != CPVS_OK) throw Error(rc, cpvs_last_error());
}

// One lazily created context per device for callers that, like the reference, have no notion of one.
inline cpvs_ctx* defaultContext(int device = 0) {
	static cpvs_ctx* ctx[16] = {nullptr};
	if (!ctx[device]) check(cpvs_ctx_create(device, &ctx[device]));
	return ctx[device];
}

#ifdef CPVS_FACADE_HAVE_GLM
inline const float* matrixData(const mat4& m) { return glm::value_ptr(m); }
#else
inline const float* matrixData(const mat4& m) { return m.m; }
#endif

}  // namespace cpvs_facade

#endif
