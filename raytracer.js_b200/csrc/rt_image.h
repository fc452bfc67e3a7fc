// rt_image.h — host decoder behind ImageTexture.  The reference decodes image files in the browser
// only (src/texture/texture_image.ts:76-136: an <img> drawn into a canvas, getImageData, RGB kept, alpha dropped);
// from Node - or any host without a DOM - the texel pool of rt_scene_desc has to be filled by a decoder of our own.
// Formats: PNG (8 bits per channel, non-interlaced: grey, grey + alpha, RGB, RGBA, palette), BMP (uncompressed 24 /
// 32 bit, bottom-up or top-down), binary PPM (P6, maxval 255).  Output: RGB8, rows top to bottom - what
// ImageTexture.image_data holds before its / 255.0.  zlib does the inflate; everything else is here.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace rt_image {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t le32(const uint8_t* p) { return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]; }
inline uint32_t le16(const uint8_t* p) { return ((uint32_t)p[1] << 8) | p[0]; }

inline bool decode_png(const uint8_t* b, size_t n, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb, std::string& err) {
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (n < 8 || memcmp(b, sig, 8) != 0) { err = "not a PNG"; return false; }
	size_t off = 8;
	int depth = 0, ctype = -1, interlace = 0;
	std::vector<uint8_t> idat, palette;
	bool have_ihdr = false, ended = false;
	while (off + 12 <= n && !ended) {
		const uint32_t len = be32(b + off);
		const uint8_t* tag = b + off + 4;
		const uint8_t* data = b + off + 8;
		if ((size_t)len > n - off - 12) { err = "PNG: truncated chunk"; return false; }
		if (!memcmp(tag, "IHDR", 4)) {
			if (len < 13) { err = "PNG: bad IHDR"; return false; }
			w = be32(data); h = be32(data + 4);
			depth = data[8]; ctype = data[9]; interlace = data[12];
			have_ihdr = true;
		} else if (!memcmp(tag, "PLTE", 4)) {
			palette.assign(data, data + len);
		} else if (!memcmp(tag, "IDAT", 4)) {
			idat.insert(idat.end(), data, data + len);
		} else if (!memcmp(tag, "IEND", 4)) {
			ended = true;
		}
		off += 12 + (size_t)len;
	}
	if (!have_ihdr || w == 0 || h == 0 || w > 65536 || h > 65536) { err = "PNG: missing or bad IHDR"; return false; }
	if (depth != 8 || interlace != 0) { err = "PNG: only 8 bits per channel, non-interlaced files are decoded"; return false; }
	int ch;
	switch (ctype) {
		case 0: ch = 1; break;
		case 2: ch = 3; break;
		case 3: ch = 1; break;
		case 4: ch = 2; break;
		case 6: ch = 4; break;
		default: err = "PNG: bad colour type"; return false;
	}
	if (ctype == 3 && palette.size() < 3) { err = "PNG: palette image without PLTE"; return false; }
	const size_t stride = (size_t)w * ch;
	std::vector<uint8_t> raw((stride + 1) * h);
	uLongf raw_len = (uLongf)raw.size();
	if (idat.empty() || uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) {
		err = "PNG: inflate failed";
		return false;
	}
	// the five scanline filters (PNG specification, section 9), in place
	std::vector<uint8_t> prev(stride, 0);
	rgb.resize((size_t)w * h * 3);
	for (uint32_t y = 0; y < h; y++) {
		uint8_t* line = raw.data() + (stride + 1) * y;
		const int f = line[0];
		uint8_t* cur = line + 1;
		for (size_t i = 0; i < stride; i++) {
			const int a = i >= (size_t)ch ? cur[i - ch] : 0, up = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
			int pred = 0;
			switch (f) {
				case 0: pred = 0; break;
				case 1: pred = a; break;
				case 2: pred = up; break;
				case 3: pred = (a + up) >> 1; break;
				case 4: {
					const int p = a + up - c, pa = p > a ? p - a : a - p, pb = p > up ? p - up : up - p, pc = p > c ? p - c : c - p;
					pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? up : c);
					break;
				}
				default: err = "PNG: bad filter type"; return false;
			}
			cur[i] = (uint8_t)(cur[i] + pred);
		}
		memcpy(prev.data(), cur, stride);
		uint8_t* out = rgb.data() + (size_t)y * w * 3;
		for (uint32_t x = 0; x < w; x++) {
			const uint8_t* px = cur + (size_t)x * ch;
			if (ctype == 2 || ctype == 6) {
				out[3 * x] = px[0]; out[3 * x + 1] = px[1]; out[3 * x + 2] = px[2];  // (alpha dropped, as the reference does)
			} else if (ctype == 3) {
				const size_t k = (size_t)px[0] * 3;
				if (k + 2 >= palette.size()) { err = "PNG: palette index out of range"; return false; }
				out[3 * x] = palette[k]; out[3 * x + 1] = palette[k + 1]; out[3 * x + 2] = palette[k + 2];
			} else {
				out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = px[0];
			}
		}
	}
	return true;
}

inline bool decode_bmp(const uint8_t* b, size_t n, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb, std::string& err) {
	if (n < 54 || b[0] != 'B' || b[1] != 'M') { err = "not a BMP"; return false; }
	const uint32_t data_off = le32(b + 10), hdr = le32(b + 14);
	if (hdr < 40) { err = "BMP: unsupported header"; return false; }
	const int32_t sw = (int32_t)le32(b + 18), sh = (int32_t)le32(b + 22);
	const uint32_t bpp = le16(b + 28), comp = le32(b + 30);
	if (sw <= 0 || sh == 0 || sw > 65536 || sh > 65536 || sh < -65536) { err = "BMP: bad size"; return false; }
	if ((bpp != 24 && bpp != 32) || (comp != 0 && !(comp == 3 && bpp == 32))) { err = "BMP: only uncompressed 24 / 32 bit files are decoded"; return false; }
	w = (uint32_t)sw;
	h = (uint32_t)(sh < 0 ? -sh : sh);
	const size_t row = ((size_t)w * (bpp / 8) + 3) & ~(size_t)3;
	if ((size_t)data_off > n || row * h > n - data_off) { err = "BMP: truncated"; return false; }
	rgb.resize((size_t)w * h * 3);
	for (uint32_t y = 0; y < h; y++) {
		const uint8_t* src = b + data_off + row * (sh < 0 ? y : h - 1 - y);  // bottom-up unless the height is negative
		uint8_t* out = rgb.data() + (size_t)y * w * 3;
		for (uint32_t x = 0; x < w; x++) {
			const uint8_t* px = src + (size_t)x * (bpp / 8);
			out[3 * x] = px[2]; out[3 * x + 1] = px[1]; out[3 * x + 2] = px[0];  // BGR(A)
		}
	}
	return true;
}

inline bool decode_ppm(const uint8_t* b, size_t n, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb, std::string& err) {
	if (n < 2 || b[0] != 'P' || b[1] != '6') { err = "not a binary PPM"; return false; }
	size_t off = 2;
	uint32_t vals[3];
	for (int k = 0; k < 3; k++) {
		for (;;) {  // whitespace and comments
			while (off < n && (b[off] == ' ' || b[off] == '\t' || b[off] == '\n' || b[off] == '\r')) off++;
			if (off < n && b[off] == '#') { while (off < n && b[off] != '\n') off++; } else break;
		}
		uint64_t v = 0;
		bool any = false;
		while (off < n && b[off] >= '0' && b[off] <= '9' && v < 1000000) { v = v * 10 + (b[off++] - '0'); any = true; }
		if (!any) { err = "PPM: bad header"; return false; }
		vals[k] = (uint32_t)v;
	}
	if (off >= n) { err = "PPM: truncated"; return false; }
	off++;  // the single whitespace byte after maxval
	w = vals[0]; h = vals[1];
	if (w == 0 || h == 0 || w > 65536 || h > 65536 || vals[2] != 255) { err = "PPM: only maxval 255 is decoded"; return false; }
	if ((size_t)w * h * 3 > n - off) { err = "PPM: truncated"; return false; }
	rgb.assign(b + off, b + off + (size_t)w * h * 3);
	return true;
}

inline bool decode(const uint8_t* b, size_t n, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb, std::string& err) {
	if (!b || n < 2) { err = "image: empty input"; return false; }
	if (b[0] == 0x89) return decode_png(b, n, w, h, rgb, err);
	if (b[0] == 'B' && b[1] == 'M') return decode_bmp(b, n, w, h, rgb, err);
	if (b[0] == 'P' && b[1] == '6') return decode_ppm(b, n, w, h, rgb, err);
	err = "image: unknown format (PNG, BMP and binary PPM are decoded)";
	return false;
}

}  // namespace rt_image
