/* Headless stand-in for <GL/glew.h>, written for this repo (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's hot-path translation units (CompressedShadow, CompressedShadowUtil,
 * MinMaxHierarchy, ShadowMap, Texture, cpvs) include <GL/glew.h> through src/cpvs.h:12 although
 * none of the DAG logic touches OpenGL. This header supplies just enough types, enumerants and
 * no-op entry points for those units to compile unmodified from /root/reference without a GL
 * installation. Every gl* call does nothing; glGenTextures hands out increasing ids.
 */
#ifndef CPVS_ORACLE_GLEW_SHIM_H
#define CPVS_ORACLE_GLEW_SHIM_H

typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
typedef int GLsizei;
typedef float GLfloat;
typedef unsigned char GLboolean;
typedef void GLvoid;
typedef unsigned int GLbitfield;
typedef char GLchar;
typedef long GLsizeiptr;
typedef long GLintptr;

enum CpvsShimGLConstants {
	GL_NO_ERROR = 0, GL_FALSE = 0, GL_TRUE = 1, GL_NONE = 0,
	GL_INVALID_ENUM = 0x0500, GL_INVALID_VALUE, GL_INVALID_OPERATION, GL_STACK_OVERFLOW,
	GL_STACK_UNDERFLOW, GL_OUT_OF_MEMORY, GL_INVALID_FRAMEBUFFER_OPERATION,
	GL_TEXTURE_1D = 0x0DE0, GL_TEXTURE_2D, GL_TEXTURE0 = 0x84C0,
	GL_DEPTH_COMPONENT = 0x1902, GL_RED, GL_GREEN, GL_BLUE, GL_ALPHA, GL_RGB, GL_RGBA,
	GL_RG = 0x8227, GL_FLOAT = 0x1406, GL_UNSIGNED_BYTE = 0x1401, GL_UNSIGNED_INT = 0x1405,
	GL_NEAREST = 0x2600, GL_LINEAR, GL_CLAMP_TO_BORDER = 0x812D, GL_CLAMP_TO_EDGE = 0x812F,
	GL_REPEAT = 0x2901,
	GL_TEXTURE_MAG_FILTER = 0x2800, GL_TEXTURE_MIN_FILTER, GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T,
	GL_TEXTURE_BORDER_COLOR = 0x1004, GL_MAX_TEXTURE_SIZE = 0x0D33,
	GL_R32F = 0x822E, GL_RG32F = 0x8230, GL_RGB32F = 0x8815, GL_RGBA32F = 0x8814,
	GL_R8 = 0x8229, GL_RGBA8 = 0x8058, GL_RGB8 = 0x8051,
	GL_DEPTH_COMPONENT16 = 0x81A5, GL_DEPTH_COMPONENT24, GL_DEPTH_COMPONENT32,
	GL_DEPTH_COMPONENT32F = 0x8CAC,
	GL_FRAMEBUFFER = 0x8D40, GL_RENDERBUFFER, GL_DRAW_FRAMEBUFFER = 0x8CA9,
	GL_READ_FRAMEBUFFER = 0x8CA8, GL_COLOR_ATTACHMENT0 = 0x8CE0, GL_DEPTH_ATTACHMENT = 0x8D00,
	GL_FRAMEBUFFER_COMPLETE = 0x8CD5,
	GL_READ_ONLY = 0x88B8, GL_WRITE_ONLY, GL_READ_WRITE,
	GL_COLOR_BUFFER_BIT = 0x4000, GL_DEPTH_BUFFER_BIT = 0x0100
};

/* Any gl* call with any argument list compiles to nothing. */
struct CpvsShimNoop {
	template <typename... A> int operator()(A...) const { return 0; }
};
#define CPVS_SHIM_NOOP(name) static const CpvsShimNoop name = CpvsShimNoop()

CPVS_SHIM_NOOP(glDeleteTextures);
CPVS_SHIM_NOOP(glBindImageTexture);
CPVS_SHIM_NOOP(glActiveTexture);
CPVS_SHIM_NOOP(glBindTexture);
CPVS_SHIM_NOOP(glTexParameteri);
CPVS_SHIM_NOOP(glTexParameterfv);
CPVS_SHIM_NOOP(glTexImage1D);
CPVS_SHIM_NOOP(glTexImage2D);
CPVS_SHIM_NOOP(glTexSubImage1D);
CPVS_SHIM_NOOP(glTexSubImage2D);
CPVS_SHIM_NOOP(glGetTexImage);
CPVS_SHIM_NOOP(glGenFramebuffers);
CPVS_SHIM_NOOP(glDeleteFramebuffers);
CPVS_SHIM_NOOP(glBindFramebuffer);
CPVS_SHIM_NOOP(glFramebufferTexture2D);
CPVS_SHIM_NOOP(glFramebufferTexture);
CPVS_SHIM_NOOP(glDrawBuffers);
CPVS_SHIM_NOOP(glDrawBuffer);
CPVS_SHIM_NOOP(glReadBuffer);
CPVS_SHIM_NOOP(glCheckFramebufferStatus);
CPVS_SHIM_NOOP(glGenRenderbuffers);
CPVS_SHIM_NOOP(glDeleteRenderbuffers);
CPVS_SHIM_NOOP(glBindRenderbuffer);
CPVS_SHIM_NOOP(glRenderbufferStorage);
CPVS_SHIM_NOOP(glFramebufferRenderbuffer);
CPVS_SHIM_NOOP(glViewport);
CPVS_SHIM_NOOP(glClear);
CPVS_SHIM_NOOP(glGetIntegerv);

static inline void glGenTextures(GLsizei n, GLuint* ids) {
	static GLuint next = 1;
	for (GLsizei i = 0; i < n; ++i) ids[i] = next++;
}
static inline GLenum glGetError() { return GL_NO_ERROR; }

#endif
