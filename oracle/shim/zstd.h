// Test-infrastructure shim (NOT product code): the five zstd prototypes that
// storage/compress/compressor_zstd.h uses, implemented as "compression
// unavailable" stubs.  The oracle never configures block compression (the raw
// vector store is MemoryOnly with an empty "compress" section), so these are
// never reached; they exist so the reference TU compiles without zstd.h.
#pragma once
#include <stddef.h>
static inline size_t ZSTD_compressBound(size_t n) { return n + 64; }
static inline size_t ZSTD_compress(void *, size_t, const void *, size_t, int) { return (size_t)-1; }
static inline unsigned ZSTD_isError(size_t code) { return code == (size_t)-1; }
static inline unsigned long long ZSTD_getDecompressedSize(const void *, size_t) { return 0; }
static inline size_t ZSTD_decompress(void *, size_t, const void *, size_t) { return (size_t)-1; }
