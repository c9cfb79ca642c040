// Build configuration for the STXXL sources vendored by the reference
// (/root/reference/stxxl). Hand-written for this Linux/gcc image so that the
// reference can be compiled by oracle/Makefile with plain g++ instead of the
// reference's own cmake build. TEST INFRASTRUCTURE ONLY (see oracle/README.md).
#ifndef STXXL_CONFIG_HEADER
#define STXXL_CONFIG_HEADER

#define STXXL_VERSION_MAJOR 1
#define STXXL_VERSION_MINOR 4
#define STXXL_VERSION_PATCH 99
#define STXXL_VERSION_STRING "1.4.99"
#define STXXL_VERSION_PHASE "prerelease/Release"

#define STXXL_HAVE_MMAP_FILE 1
#define STXXL_HAVE_LINUXAIO_FILE 1
#define STXXL_POSIX_THREADS 1
#define STXXL_HAVE_CXX11_RANGE_FOR_LOOP 1
#define STXXL_HAVE_SYNC_ADD_AND_FETCH 1
#define STXXL_STD_RANDOM 1
#define STXXL_HAVE_MLOCK_PROTO 1

#endif // !STXXL_CONFIG_HEADER
