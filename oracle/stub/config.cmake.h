/* Stub build configuration for compiling the reference's libgimli sources
 * (from /root/reference/core/src, never copied here) without cmake.
 * Mirrors the switches of core/config.cmake.h.in with every optional
 * third-party package turned off. Test infrastructure only (oracle/_ref). */
#ifndef LIBGIMLI_CONFIG__H
#define LIBGIMLI_CONFIG__H
#define PACKAGE_NAME "libgimli"
#define PACKAGE_BUGREPORT "none"
#define PACKAGE_AUTHORS "gimli-org"
#define LIBGIMLI_VERSION_MAJOR 1
#define LIBGIMLI_VERSION_MINOR 6
#define LIBGIMLI_VERSION_PATCH 0
#define LIBGIMLI_VERSION "1.6.0"
#define PACKAGE_VERSION "1.6.0-oracle"
#define SRC_DIR "/root/reference"
#define HAVE_BOOST_INTERPROCESS_MANAGED_SHARED_MEMORY_HPP 0
#define BOOST_THREAD_FOUND 0
#define BOOST_BIND_FOUND 0
#define TRIANGLE_FOUND 0
#define CHOLMOD_FOUND 0
#define OPENBLAS_FOUND 0
#define OPENBLAS_CBLAS_FOUND 0
#define CONDA_BUILD 0
#define USE_IPC 0
#define READPROC_FOUND 0
#endif
