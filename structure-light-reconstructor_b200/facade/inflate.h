// inflate.h — zlib-stream decoder of the PNG ingest path (see inflate.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace duke {

// Inflates a zlib stream (RFC 1950: 2-byte header, deflate data, Adler-32) whose decompressed size is known — a PNG's
// concatenated IDAT payload is height x (1 + row bytes).  false (with *err) on malformed input, a size mismatch or a
// checksum mismatch; never reads or writes outside [in, in + in_len) / [out, out + out_len).
bool zlib_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, std::string *err = nullptr);

// The deflate data alone (RFC 1951).  *in_used = bytes of input consumed.
bool raw_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, size_t *in_used, std::string *err = nullptr);

}  // namespace duke
