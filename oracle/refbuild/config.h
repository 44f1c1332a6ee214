/* Hand-written Linux x86-64 config.h used ONLY to compile the unmodified reference
 * sources (read in place from /root/reference/src) into oracle/_ref/.  It replaces the
 * autoconf-generated header (the checkout's own src/config.h is a macOS/arm64 leftover
 * and src/Makefile.in is missing, so the reference's build system cannot run here).
 * This file is test infrastructure, not product code. */
#ifndef GMAPDP_ORACLE_REF_CONFIG_H
#define GMAPDP_ORACLE_REF_CONFIG_H
#define HAVE_BUILTIN_CLZ 1
#define HAVE_BUILTIN_CLZLL 1
#define HAVE_BUILTIN_CTZ 1
#define HAVE_BUILTIN_CTZLL 1
#define HAVE_BUILTIN_POPCOUNT 1
#define HAVE_CADDR_T 1
#define HAVE_CEIL 1
#define HAVE_DIRENT_H 1
#define HAVE_FCNTL_H 1
#define HAVE_FLOOR 1
#define HAVE_FSEEKO 1
#define HAVE_INDEX 1
#define HAVE_INLINE 1
#define HAVE_INTTYPES_H 1
#define HAVE_LIBM 1
#define HAVE_LIMITS_H 1
#define HAVE_LOG 1
#define HAVE_MADVISE 1
#define HAVE_MADVISE_MADV_DONTNEED 1
#define HAVE_MADVISE_MADV_RANDOM 1
#define HAVE_MADVISE_MADV_SEQUENTIAL 1
#define HAVE_MADVISE_MADV_WILLNEED 1
#define HAVE_MEMCPY 1
#define HAVE_MEMMOVE 1
#define HAVE_MEMSET 1
#define HAVE_MMAP 1
#define HAVE_MMAP_MAP_FILE 1
#define HAVE_MMAP_MAP_PRIVATE 1
#define HAVE_MMAP_MAP_SHARED 1
#define HAVE_MM_EXTRACT_EPI64 1
#define HAVE_MM_POPCNT 1
#define HAVE_MM_POPCNT_U64 1
#define HAVE_MUNMAP 1
#define HAVE_POPCNT 1
#define HAVE_POW 1
#define HAVE_PTHREAD 1
#define HAVE_RINT 1
#define HAVE_SEMCTL 1
#define HAVE_SEMGET 1
#define HAVE_SEMOP 1
#define HAVE_SHMAT 1
#define HAVE_SHMCTL 1
#define HAVE_SHMDT 1
#define HAVE_SHMGET 1
#define HAVE_SIGACTION 1
#define HAVE_STAT64 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_STRTOUL 1
#define HAVE_STRUCT_STAT64 1
#define HAVE_SYSCONF 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_ZLIB 1
#define HAVE_ZLIB_GZBUFFER 1
#define PACKAGE "gmap"
#define PACKAGE_BUGREPORT "n/a"
#define PACKAGE_NAME "gmap"
#define PACKAGE_STRING "gmap 2024-02-22"
#define PACKAGE_TARNAME "gmap"
#define PACKAGE_URL ""
#define PACKAGE_VERSION "2024-02-22"
#define PAGESIZE_VIA_SYSCONF 1
#define SIZEOF_OFF_T 8
#define SIZEOF_UNSIGNED_LONG 8
#define SIZEOF_UNSIGNED_LONG_LONG 8
#define SSE2_SLLI_CONST_IMM8 1
#define STDC_HEADERS 1
#define USE_FOPEN_BINARY 1
#define USE_FOPEN_TEXT 1
#define USE_INTEL_INTRINSICS 1
#define VERSION "2024-02-22"
#define TARGET "x86_64-pc-linux-gnu"
#define GMAPDB "/tmp/gmapdb"
#endif
