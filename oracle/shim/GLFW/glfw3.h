/* Headless stand-in for <GLFW/glfw3.h> (see ../GL/glew.h). The hot-path units need nothing from it. */
#ifndef CPVS_ORACLE_GLFW_SHIM_H
#define CPVS_ORACLE_GLFW_SHIM_H
#endif
