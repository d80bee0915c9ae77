// See exr_codec.h. File layout (OpenEXR 2.x, single-part scan-line image), all integers little-endian:
//   magic 0x01312f76 | version byte 2 + flag bits (0x200 tiled, 0x800 deep, 0x1000 multi-part)
//   attributes: name\0 type\0 int32 size, payload ... terminated by an empty name
//   offset table: one uint64 per chunk | chunks: int32 y, int32 byteCount, payload
// A chunk holds 1 (NONE, RLE, ZIPS), 16 (ZIP) or 32 (PIZ) scan lines; decompressed it is, line by line, the channels
// in alphabetical order, `width` little-endian values each. A payload as long as the raw block is stored raw.
#include "exr_codec.h"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>

namespace bcdio
{

// ------------------------------------------------------------------------------------------------ half <-> float
float halfToFloat(unsigned short h)
{
	const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
	if(exp == 0)
	{
		if(man == 0)
			bits = sign;
		else
		{
			int e = -1; // normalise the subnormal
			do { ++e; man <<= 1; } while(!(man & 0x400u));
			bits = sign | (uint32_t)(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
		}
	}
	else if(exp == 31)
		bits = sign | 0x7f800000u | man << 13;
	else
		bits = sign | (exp + 127 - 15) << 23 | man << 13;
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

unsigned short floatToHalf(float f)
{
	uint32_t x;
	memcpy(&x, &f, 4);
	const uint32_t sign = (x >> 16) & 0x8000u;
	const int e = (int)((x >> 23) & 0xffu) - 127 + 15;
	uint32_t m = x & 0x007fffffu;
	if(e <= 0)
	{
		if(e < -10)
			return (unsigned short)sign; // rounds to zero
		m |= 0x00800000u; // subnormal half: shift the significand, round to nearest even
		const int t = 14 - e;
		const uint32_t a = (1u << (t - 1)) - 1u, b = (m >> t) & 1u;
		return (unsigned short)(sign | ((m + a + b) >> t));
	}
	if(e == 0xff - 127 + 15)
	{
		if(m == 0)
			return (unsigned short)(sign | 0x7c00u); // infinity
		m >>= 13;
		return (unsigned short)(sign | 0x7c00u | m | (m == 0)); // NaN stays a NaN
	}
	m = m + 0x00000fffu + ((m >> 13) & 1u); // round to nearest even
	int ee = e;
	if(m & 0x00800000u)
	{
		m = 0;
		ee += 1;
	}
	if(ee > 30)
		return (unsigned short)(sign | 0x7c00u); // overflow
	return (unsigned short)(sign | (uint32_t)ee << 10 | (m >> 13));
}

namespace
{

struct Reader
{
	const uint8_t* p;
	const uint8_t* end;
	bool ok;
	Reader(const uint8_t* b, size_t n) : p(b), end(b + n), ok(true) {}
	bool need(size_t n)
	{
		if((size_t)(end - p) < n) ok = false;
		return ok;
	}
	uint8_t u8() { if(!need(1)) return 0; return *p++; }
	uint32_t u32()
	{
		if(!need(4)) return 0;
		const uint32_t v = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
		p += 4;
		return v;
	}
	int32_t i32() { return (int32_t)u32(); }
	uint64_t u64()
	{
		const uint64_t lo = u32(), hi = u32();
		return lo | hi << 32;
	}
	std::string str()
	{
		std::string s;
		while(need(1) && *p)
			s.push_back((char)*p++);
		if(ok) ++p;
		return s;
	}
};

int linesPerChunk(int compression)
{
	switch(compression)
	{
		case EXR_NONE: case EXR_RLE: case EXR_ZIPS: return 1;
		case EXR_ZIP: return 16;
		case EXR_PIZ: return 32;
		default: return 0;
	}
}

int bytesOf(int type) { return type == EXR_HALF ? 2 : 4; }

// ---- the byte predictor + even/odd split shared by RLE and ZIP (decoding direction)
void undoPredictorAndSplit(std::vector<uint8_t>& t, uint8_t* out)
{
	const size_t n = t.size();
	for(size_t i = 1; i < n; ++i)
		t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
	const size_t half = (n + 1) / 2;
	for(size_t i = 0, a = 0, b = half; i < n;)
	{
		out[i++] = t[a++];
		if(i < n) out[i++] = t[b++];
	}
}

void applySplitAndPredictor(const uint8_t* raw, size_t n, std::vector<uint8_t>& t)
{
	t.resize(n);
	const size_t half = (n + 1) / 2;
	for(size_t i = 0, a = 0, b = half; i < n;)
	{
		t[a++] = raw[i++];
		if(i < n) t[b++] = raw[i++];
	}
	int prev = n ? t[0] : 0;
	for(size_t i = 1; i < n; ++i)
	{
		const int cur = t[i];
		t[i] = (uint8_t)(cur - prev + 128 + 256);
		prev = cur;
	}
}

bool unzipBlock(const uint8_t* src, size_t n, size_t rawSize, uint8_t* out)
{
	std::vector<uint8_t> t(rawSize);
	uLongf got = (uLongf)rawSize;
	if(uncompress(t.data(), &got, src, (uLong)n) != Z_OK || got != rawSize)
		return false;
	undoPredictorAndSplit(t, out);
	return true;
}

bool unrleBlock(const uint8_t* src, size_t n, size_t rawSize, uint8_t* out)
{
	std::vector<uint8_t> t;
	t.reserve(rawSize);
	size_t i = 0;
	while(i < n)
	{
		const int c = (int8_t)src[i++];
		if(c < 0)
		{
			const size_t cnt = (size_t)(-c);
			if(i + cnt > n || t.size() + cnt > rawSize) return false;
			t.insert(t.end(), src + i, src + i + cnt);
			i += cnt;
		}
		else
		{
			const size_t cnt = (size_t)c + 1;
			if(i >= n || t.size() + cnt > rawSize) return false;
			t.insert(t.end(), cnt, src[i++]);
		}
	}
	if(t.size() != rawSize) return false;
	undoPredictorAndSplit(t, out);
	return true;
}

// ------------------------------------------------------------------------------------------------------- PIZ
// Huffman stream (ImfHuf): 20-byte header {minSymbol, maxSymbol, -, bitCount, -}, the code lengths of the symbols
// minSymbol..maxSymbol packed 6 bits each (59..62: run of 2..5 zero lengths, 63: 8 more bits = run - 6), then the
// MSB-first code stream. Codes are canonical: within a length, by ascending symbol; the first code of length l-1 is
// (first(l) + count(l)) >> 1, so longer codes are numerically smaller. Symbol maxSymbol is the run-length escape:
// the next 8 bits repeat the previous output value.
struct BitIn
{
	const uint8_t* p;
	const uint8_t* end;
	uint64_t acc;
	int n;
	BitIn(const uint8_t* b, const uint8_t* e) : p(b), end(e), acc(0), n(0) {}
	void fill(int want) // make `want` (<= 56) bits available, zero-padded past the end
	{
		while(n < want)
		{
			acc = (acc << 8) | (p < end ? *p : 0);
			++p;
			n += 8;
		}
	}
	uint32_t peek(int k) { fill(k); return (uint32_t)((acc >> (n - k)) & ((1ull << k) - 1)); }
	uint64_t peek64(int k) { fill(k); return (acc >> (n - k)) & ((1ull << k) - 1); }
	void drop(int k) { n -= k; }
	uint32_t get(int k) { const uint32_t v = peek(k); drop(k); return v; }
};

bool hufDecode(const uint8_t* comp, size_t nComp, uint16_t* raw, size_t nRaw, std::string& err)
{
	if(nComp == 0)
	{
		if(nRaw != 0) err = "PIZ: empty Huffman block";
		return nRaw == 0;
	}
	if(nComp < 20) { err = "PIZ: truncated Huffman header"; return false; }
	Reader hd(comp, 20);
	const uint32_t im = hd.u32(), iM = hd.u32();
	hd.u32();
	const uint32_t nBits = hd.u32();
	const uint32_t NSYM = 65537;
	if(im >= NSYM || iM >= NSYM || im > iM) { err = "PIZ: bad Huffman symbol range"; return false; }
	std::vector<uint8_t> len(NSYM, 0);
	BitIn tb(comp + 20, comp + nComp);
	for(uint32_t s = im; s <= iM; ++s)
	{
		const uint32_t l = tb.get(6);
		if(l == 63 || l >= 59)
		{
			const uint32_t run = l == 63 ? tb.get(8) + 6 : l - 59 + 2;
			if(s + run > iM + 1) { err = "PIZ: Huffman table too long"; return false; }
			s += run - 1;
		}
		else
			len[s] = (uint8_t)l;
	}
	// the table reader pulls whole bytes on demand: the code stream starts after the last byte it touched
	const uint8_t* data = tb.p - (tb.n / 8);
	if(data > comp + nComp || (uint64_t)nBits > 8ull * (uint64_t)(comp + nComp - data)) { err = "PIZ: bad Huffman bit count"; return false; }
	uint64_t count[59] = {0}, first[59] = {0};
	for(uint32_t s = 0; s < NSYM; ++s)
		count[len[s]]++;
	{
		uint64_t c = 0;
		for(int l = 58; l > 0; --l)
		{
			first[l] = c;
			c = (c + count[l]) >> 1;
		}
	}
	std::vector<uint32_t> start(60, 0), bySym;
	for(int l = 1; l <= 58; ++l)
		start[l + 1] = start[l] + (uint32_t)count[l];
	bySym.resize(start[59]);
	{
		std::vector<uint32_t> fillPos(start.begin(), start.end());
		for(uint32_t s = 0; s < NSYM; ++s)
			if(len[s]) bySym[fillPos[len[s]]++] = s;
	}
	const int FAST = 12;
	std::vector<uint32_t> fast(1u << FAST, 0); // symbol << 6 | length, 0 = longer code
	for(int l = 1; l <= FAST; ++l)
		for(uint64_t k = 0; k < count[l]; ++k)
		{
			const uint64_t code = first[l] + k;
			if(code >> l) { err = "PIZ: inconsistent Huffman lengths"; return false; }
			const uint32_t lo = (uint32_t)(code << (FAST - l)), entry = bySym[start[l] + (uint32_t)k] << 6 | (uint32_t)l;
			for(uint32_t j = 0; j < (1u << (FAST - l)); ++j)
				fast[lo + j] = entry;
		}
	BitIn in(data, data + (nBits + 7) / 8);
	uint64_t used = 0;
	size_t o = 0;
	while(o < nRaw)
	{
		uint32_t sym = 0;
		int l = 0;
		const uint32_t e = fast[in.peek(FAST)];
		if(e)
		{
			sym = e >> 6;
			l = (int)(e & 63);
		}
		else
		{
			for(l = FAST + 1; l <= 56; ++l)
			{
				if(!count[l]) continue;
				const uint64_t code = in.peek64(l);
				if(code >= first[l] && code - first[l] < count[l])
				{
					sym = bySym[start[l] + (uint32_t)(code - first[l])];
					break;
				}
			}
			if(l > 56) { err = "PIZ: invalid Huffman code"; return false; }
		}
		in.drop(l);
		used += (uint64_t)l;
		if(used > nBits) { err = "PIZ: Huffman stream too short"; return false; }
		if(sym == iM)
		{
			const uint32_t run = in.get(8);
			used += 8;
			if(used > nBits || o == 0 || o + run > nRaw) { err = "PIZ: bad run in Huffman stream"; return false; }
			const uint16_t v = raw[o - 1];
			for(uint32_t k = 0; k < run; ++k)
				raw[o++] = v;
		}
		else
			raw[o++] = (uint16_t)sym;
	}
	return true;
}

// inverse of the 2 x 2 Haar-like lifting step of ImfWav (14-bit variant: plain shorts; 16-bit: modulo 2^16)
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
	const int ls = (int16_t)l, hi = (int16_t)h;
	const int ai = ls + (hi & 1) + (hi >> 1);
	a = (uint16_t)(int16_t)ai;
	b = (uint16_t)(int16_t)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
	const int m = l, d = h;
	const int bb = (m - (d >> 1)) & 0xffff;
	a = (uint16_t)((d + bb - 0x8000) & 0xffff);
	b = (uint16_t)bb;
}

void waveletDecode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t maxValue)
{
	const bool w14 = maxValue < (1 << 14);
	const int n = nx > ny ? ny : nx;
	int p = 1;
	while(p <= n) p <<= 1;
	p >>= 1;
	int p2 = p;
	p >>= 1;
	auto dec = [w14](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) { if(w14) wdec14(l, h, a, b); else wdec16(l, h, a, b); };
	for(; p >= 1; p2 = p, p >>= 1)
	{
		const long oy1 = (long)oy * p, oy2 = (long)oy * p2, ox1 = (long)ox * p, ox2 = (long)ox * p2;
		uint16_t* py = in;
		uint16_t* const ey = in + (long)oy * (ny - p2);
		for(; py <= ey; py += oy2)
		{
			uint16_t* px = py;
			uint16_t* const ex = py + (long)ox * (nx - p2);
			for(; px <= ex; px += ox2)
			{
				uint16_t* p01 = px + ox1;
				uint16_t* p10 = px + oy1;
				uint16_t* p11 = p10 + ox1;
				uint16_t i00, i01, i10, i11;
				dec(*px, *p10, i00, i10);
				dec(*p01, *p11, i01, i11);
				dec(i00, i01, *px, *p01);
				dec(i10, i11, *p10, *p11);
			}
			if(nx & p) // odd column at this level: vertical step only
			{
				uint16_t* p10 = px + oy1;
				uint16_t i00;
				dec(*px, *p10, i00, *p10);
				*px = i00;
			}
		}
		if(ny & p) // odd line: horizontal step only
		{
			uint16_t* px = py;
			uint16_t* const ex = py + (long)ox * (nx - p2);
			for(; px <= ex; px += ox2)
			{
				uint16_t* p01 = px + ox1;
				uint16_t i00;
				dec(*px, *p01, i00, *p01);
				*px = i00;
			}
		}
	}
}

bool unpizBlock(const uint8_t* src, size_t n, int width, int lines, const std::vector<int>& types, uint8_t* out,
		std::string& err)
{
	const size_t nch = types.size();
	std::vector<size_t> startOf(nch + 1, 0);
	for(size_t c = 0; c < nch; ++c)
		startOf[c + 1] = startOf[c] + (size_t)width * lines * (bytesOf(types[c]) / 2);
	const size_t total = startOf[nch];
	Reader r(src, n);
	const uint32_t lo = (uint32_t)r.u8() | (uint32_t)r.u8() << 8, hi = (uint32_t)r.u8() | (uint32_t)r.u8() << 8;
	if(!r.ok || hi >= 8192) { err = "PIZ: bad bitmap range"; return false; }
	std::vector<uint8_t> bitmap(8192, 0);
	if(lo <= hi)
	{
		if(!r.need(hi - lo + 1)) { err = "PIZ: truncated bitmap"; return false; }
		memcpy(bitmap.data() + lo, r.p, hi - lo + 1);
		r.p += hi - lo + 1;
	}
	std::vector<uint16_t> lut(65536, 0);
	uint32_t k = 0;
	for(uint32_t i = 0; i < 65536; ++i)
		if(i == 0 || (bitmap[i >> 3] & (1u << (i & 7))))
			lut[k++] = (uint16_t)i;
	const uint16_t maxValue = (uint16_t)(k - 1);
	const int32_t length = r.i32();
	if(!r.ok || length < 0 || !r.need((size_t)length)) { err = "PIZ: truncated Huffman block"; return false; }
	std::vector<uint16_t> tmp(total);
	if(!hufDecode(r.p, (size_t)length, tmp.data(), total, err))
		return false;
	for(size_t c = 0; c < nch; ++c)
	{
		const int size = bytesOf(types[c]) / 2;
		for(int j = 0; j < size; ++j)
			waveletDecode(tmp.data() + startOf[c] + j, width, size, lines, width * size, maxValue);
	}
	for(size_t i = 0; i < total; ++i)
		tmp[i] = lut[tmp[i]];
	// back to the scan-line interleaved byte layout
	std::vector<size_t> cur(startOf.begin(), startOf.end() - 1);
	uint8_t* o = out;
	for(int y = 0; y < lines; ++y)
		for(size_t c = 0; c < nch; ++c)
		{
			const size_t cnt = (size_t)width * (bytesOf(types[c]) / 2);
			for(size_t i = 0; i < cnt; ++i)
			{
				const uint16_t v = tmp[cur[c]++];
				*o++ = (uint8_t)(v & 0xff);
				*o++ = (uint8_t)(v >> 8);
			}
		}
	return true;
}

// ---- PIZ, encoding direction (mirror images of the decoders above)
struct BitOut
{
	std::vector<uint8_t>& b;
	uint64_t acc;
	int n;
	uint64_t total;
	explicit BitOut(std::vector<uint8_t>& buf) : b(buf), acc(0), n(0), total(0) {}
	void put(int k, uint64_t v) // k <= 58 bits, MSB first
	{
		total += (uint64_t)k;
		while(k > 0)
		{
			const int take = std::min(k, 32);
			acc = (acc << take) | ((v >> (k - take)) & ((1ull << take) - 1));
			n += take;
			k -= take;
			while(n >= 8)
			{
				b.push_back((uint8_t)(acc >> (n - 8)));
				n -= 8;
			}
		}
	}
	void flush()
	{
		if(n > 0) b.push_back((uint8_t)(acc << (8 - n)));
		n = 0;
	}
};

// Huffman code lengths by pairwise merging of the two rarest subtrees (a heap over subtree weights; every merge adds one
// bit to the symbols of both subtrees, kept as linked lists)
void huffmanLengths(const std::vector<uint64_t>& freq, std::vector<uint8_t>& len)
{
	const size_t n = freq.size();
	len.assign(n, 0);
	std::vector<uint32_t> syms;
	for(uint32_t i = 0; i < n; ++i)
		if(freq[i]) syms.push_back(i);
	if(syms.size() == 1) { len[syms[0]] = 1; return; }
	std::vector<uint32_t> next(n), tail(n);
	std::vector<uint64_t> weight(n);
	for(uint32_t i : syms) { next[i] = i; tail[i] = i; weight[i] = freq[i]; } // a list ends when next == itself
	auto cmp = [&](uint32_t a, uint32_t b) { return weight[a] > weight[b] || (weight[a] == weight[b] && a > b); };
	std::vector<uint32_t> heap(syms);
	std::make_heap(heap.begin(), heap.end(), cmp);
	while(heap.size() > 1)
	{
		std::pop_heap(heap.begin(), heap.end(), cmp);
		const uint32_t a = heap.back();
		heap.pop_back();
		std::pop_heap(heap.begin(), heap.end(), cmp);
		const uint32_t b = heap.back();
		heap.pop_back();
		for(uint32_t head : {a, b})
			for(uint32_t j = head;; j = next[j])
			{
				++len[j];
				if(next[j] == j) break;
			}
		next[tail[a]] = b; // concatenate b's list behind a's
		tail[a] = tail[b];
		weight[a] += weight[b];
		heap.push_back(a);
		std::push_heap(heap.begin(), heap.end(), cmp);
	}
}

void hufEncode(const uint16_t* raw, size_t nRaw, std::vector<uint8_t>& out)
{
	if(nRaw == 0) return;
	const uint32_t NSYM = 65537;
	std::vector<uint64_t> freq(NSYM, 0);
	for(size_t i = 0; i < nRaw; ++i)
		freq[raw[i]]++;
	uint32_t im = 0, iM = 0;
	while(!freq[im]) ++im;
	for(uint32_t i = im; i < 65536; ++i)
		if(freq[i]) iM = i;
	++iM; // the run-length escape is the symbol after the largest one in use
	freq[iM] = 1;
	std::vector<uint8_t> len;
	huffmanLengths(freq, len);
	uint64_t count[59] = {0}, first[59] = {0};
	for(uint32_t s = im; s <= iM; ++s)
		count[len[s]]++;
	{
		uint64_t c = 0;
		for(int l = 58; l > 0; --l)
		{
			first[l] = c;
			c = (c + count[l]) >> 1;
		}
	}
	std::vector<uint64_t> code(NSYM, 0);
	{
		uint64_t nextCode[59];
		for(int l = 0; l < 59; ++l) nextCode[l] = first[l];
		for(uint32_t s = im; s <= iM; ++s)
			if(len[s]) code[s] = nextCode[len[s]]++;
	}
	const size_t headerAt = out.size();
	out.resize(headerAt + 20, 0);
	// packed code lengths
	const size_t tableAt = out.size();
	{
		BitOut tb(out);
		for(uint32_t s = im; s <= iM; ++s)
		{
			if(len[s] == 0)
			{
				uint32_t run = 1;
				while(s < iM && run < 255 + 6 && len[s + 1] == 0) { ++s; ++run; }
				if(run >= 6) { tb.put(6, 63); tb.put(8, run - 6); continue; }
				if(run >= 2) { tb.put(6, 59 + run - 2); continue; }
			}
			tb.put(6, len[s]);
		}
		tb.flush();
	}
	const uint32_t tableLength = (uint32_t)(out.size() - tableAt);
	BitOut db(out);
	auto send = [&](uint32_t s, uint32_t extra) // symbol s repeated extra + 1 times
	{
		if((uint64_t)len[s] + len[iM] + 8 < (uint64_t)len[s] * extra)
		{
			db.put(len[s], code[s]);
			db.put(len[iM], code[iM]);
			db.put(8, extra);
		}
		else
			for(uint32_t k = 0; k <= extra; ++k)
				db.put(len[s], code[s]);
	};
	uint32_t s = raw[0], extra = 0;
	for(size_t i = 1; i < nRaw; ++i)
	{
		if(raw[i] == s && extra < 255)
			++extra;
		else
		{
			send(s, extra);
			extra = 0;
		}
		s = raw[i];
	}
	send(s, extra);
	db.flush();
	const uint32_t hdr[5] = {im, iM, tableLength, (uint32_t)db.total, 0u};
	for(int k = 0; k < 5; ++k)
		for(int b8 = 0; b8 < 4; ++b8)
			out[headerAt + 4 * k + b8] = (uint8_t)(hdr[k] >> (8 * b8));
}

inline void wenc14(uint16_t a, uint16_t b, uint16_t& l, uint16_t& h)
{
	const int as = (int16_t)a, bs = (int16_t)b;
	l = (uint16_t)(int16_t)((as + bs) >> 1);
	h = (uint16_t)(int16_t)(as - bs);
}
inline void wenc16(uint16_t a, uint16_t b, uint16_t& l, uint16_t& h)
{
	const int ao = (a + 0x8000) & 0xffff;
	int m = (ao + b) >> 1;
	int d = ao - b;
	if(d < 0) m = (m + 0x8000) & 0xffff;
	d &= 0xffff;
	l = (uint16_t)m;
	h = (uint16_t)d;
}

void waveletEncode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t maxValue)
{
	const bool w14 = maxValue < (1 << 14);
	const int n = nx > ny ? ny : nx;
	auto enc = [w14](uint16_t a, uint16_t b, uint16_t& l, uint16_t& h) { if(w14) wenc14(a, b, l, h); else wenc16(a, b, l, h); };
	for(int p = 1, p2 = 2; p2 <= n; p = p2, p2 <<= 1)
	{
		const long oy1 = (long)oy * p, oy2 = (long)oy * p2, ox1 = (long)ox * p, ox2 = (long)ox * p2;
		uint16_t* py = in;
		uint16_t* const ey = in + (long)oy * (ny - p2);
		for(; py <= ey; py += oy2)
		{
			uint16_t* px = py;
			uint16_t* const ex = py + (long)ox * (nx - p2);
			for(; px <= ex; px += ox2)
			{
				uint16_t* p01 = px + ox1;
				uint16_t* p10 = px + oy1;
				uint16_t* p11 = p10 + ox1;
				uint16_t i00, i01, i10, i11;
				enc(*px, *p01, i00, i01);
				enc(*p10, *p11, i10, i11);
				enc(i00, i10, *px, *p10);
				enc(i01, i11, *p01, *p11);
			}
			if(nx & p)
			{
				uint16_t* p10 = px + oy1;
				uint16_t i00;
				enc(*px, *p10, i00, *p10);
				*px = i00;
			}
		}
		if(ny & p)
		{
			uint16_t* px = py;
			uint16_t* const ex = py + (long)ox * (nx - p2);
			for(; px <= ex; px += ox2)
			{
				uint16_t* p01 = px + ox1;
				uint16_t i00;
				enc(*px, *p01, i00, *p01);
				*px = i00;
			}
		}
	}
}

// raw: scan-line interleaved little-endian block (lines x channels x width); sizes[c] = bytes per value of channel c
void pizBlock(const uint8_t* raw, int width, int lines, const std::vector<int>& sizes, std::vector<uint8_t>& out)
{
	const size_t nch = sizes.size();
	std::vector<size_t> startOf(nch + 1, 0);
	for(size_t c = 0; c < nch; ++c)
		startOf[c + 1] = startOf[c] + (size_t)width * lines * (sizes[c] / 2);
	const size_t total = startOf[nch];
	std::vector<uint16_t> tmp(total);
	{
		std::vector<size_t> cur(startOf.begin(), startOf.end() - 1);
		const uint8_t* s = raw;
		for(int y = 0; y < lines; ++y)
			for(size_t c = 0; c < nch; ++c)
				for(size_t i = 0, cnt = (size_t)width * (sizes[c] / 2); i < cnt; ++i, s += 2)
					tmp[cur[c]++] = (uint16_t)(s[0] | s[1] << 8);
	}
	std::vector<uint8_t> bitmap(8192, 0);
	for(size_t i = 0; i < total; ++i)
		bitmap[tmp[i] >> 3] |= (uint8_t)(1u << (tmp[i] & 7));
	bitmap[0] &= (uint8_t)~1u; // zero is always in the table, never in the bitmap
	int lo = 8191, hi = 0;
	for(int i = 0; i < 8192; ++i)
		if(bitmap[i]) { lo = std::min(lo, i); hi = std::max(hi, i); }
	std::vector<uint16_t> lut(65536, 0);
	uint32_t k = 0;
	for(uint32_t i = 0; i < 65536; ++i)
		if(i == 0 || (bitmap[i >> 3] & (1u << (i & 7))))
			lut[i] = (uint16_t)k++;
	const uint16_t maxValue = (uint16_t)(k - 1);
	for(size_t i = 0; i < total; ++i)
		tmp[i] = lut[tmp[i]];
	for(size_t c = 0; c < nch; ++c)
		for(int j = 0, size = sizes[c] / 2; j < size; ++j)
			waveletEncode(tmp.data() + startOf[c] + j, width, size, lines, width * size, maxValue);
	out.clear();
	out.push_back((uint8_t)(lo & 0xff)); out.push_back((uint8_t)(lo >> 8));
	out.push_back((uint8_t)(hi & 0xff)); out.push_back((uint8_t)(hi >> 8));
	if(lo <= hi)
		out.insert(out.end(), bitmap.begin() + lo, bitmap.begin() + hi + 1);
	const size_t lengthAt = out.size();
	out.resize(lengthAt + 4, 0);
	hufEncode(tmp.data(), total, out);
	const uint32_t length = (uint32_t)(out.size() - lengthAt - 4);
	for(int b8 = 0; b8 < 4; ++b8)
		out[lengthAt + b8] = (uint8_t)(length >> (8 * b8));
}

bool readWholeFile(const char* path, std::vector<uint8_t>& buf)
{
	FILE* f = fopen(path, "rb");
	if(!f) return false;
	fseek(f, 0, SEEK_END);
	const long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	if(n < 0) { fclose(f); return false; }
	buf.resize((size_t)n);
	const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
	fclose(f);
	return got == (size_t)n;
}

} // namespace

// ------------------------------------------------------------------------------------------------------- read
static bool exrOpenBody(const char* path, ExrReader& rd, std::string& err)
{
	std::vector<uint8_t>& buf = rd.file;
	ExrFile& out = rd.info;
	if(!readWholeFile(path, buf)) { err = std::string("cannot read '") + path + "'"; return false; }
	Reader r(buf.data(), buf.size());
	if(r.u32() != 20000630u || !r.ok) { err = std::string("'") + path + "' is not an OpenEXR file"; return false; }
	const uint32_t version = r.u32();
	if((version & 0xff) != 2) { err = "unsupported OpenEXR version"; return false; }
	if(version & 0x200) { err = "tiled OpenEXR files are not supported"; return false; }
	if(version & 0x800) { err = "deep OpenEXR files are not supported"; return false; }
	if(version & 0x1000) { err = "multi-part OpenEXR files are not supported"; return false; }
	bool haveCh = false, haveDw = false, haveComp = false;
	int maxX = 0, maxY = 0;
	out = ExrFile();
	while(r.ok)
	{
		const std::string name = r.str();
		if(name.empty()) break;
		const std::string type = r.str();
		const int32_t size = r.i32();
		if(!r.ok || size < 0 || !r.need((size_t)size)) { err = "truncated header"; return false; }
		Reader a(r.p, (size_t)size);
		r.p += size;
		if(name == "channels" && type == "chlist")
		{
			while(a.ok)
			{
				const std::string ch = a.str();
				if(ch.empty()) break;
				const int32_t pt = a.i32();
				a.u8(); a.u8(); a.u8(); a.u8(); // pLinear + reserved
				const int32_t xs = a.i32(), ys = a.i32();
				if(!a.ok) break;
				if(pt < 0 || pt > 2) { err = "unknown pixel type in channel list"; return false; }
				if(xs != 1 || ys != 1) { err = "sub-sampled channels are not supported"; return false; }
				out.names.push_back(ch);
				out.types.push_back(pt);
			}
			haveCh = true;
		}
		else if(name == "compression" && type == "compression")
		{
			out.compression = a.u8();
			haveComp = true;
		}
		else if(name == "dataWindow" && type == "box2i")
		{
			out.minX = a.i32(); out.minY = a.i32(); maxX = a.i32(); maxY = a.i32();
			haveDw = true;
		}
	}
	if(!r.ok || !haveCh || !haveDw || !haveComp) { err = "incomplete OpenEXR header"; return false; }
	const int lpc = linesPerChunk(out.compression);
	if(lpc == 0)
	{
		static const char* names[] = {"NONE", "RLE", "ZIPS", "ZIP", "PIZ", "PXR24", "B44", "B44A", "DWAA", "DWAB"};
		err = std::string("unsupported compression ") + (out.compression < 10 ? names[out.compression] : "?")
			+ " (supported: NONE, RLE, ZIPS, ZIP, PIZ)";
		return false;
	}
	const long w = (long)maxX - out.minX + 1, h = (long)maxY - out.minY + 1;
	if(w <= 0 || h <= 0 || w > (1 << 20) || h > (1 << 20)) { err = "bad data window"; return false; }
	out.width = (int)w;
	out.height = (int)h;
	const size_t nch = out.names.size();
	// every chunk costs at least its 8-byte header and its table entry: a header whose window cannot fit the file is refused
	// before anything is allocated
	if(nch == 0 || nch > 100000 || (uint64_t)((h + lpc - 1) / lpc) * 16 > (uint64_t)buf.size()) { err = "data window does not fit the file"; return false; }
	size_t lineBytes = 0;
	for(size_t c = 0; c < nch; ++c)
		lineBytes += (size_t)w * bytesOf(out.types[c]);
	rd.lineBytes = lineBytes;
	const int nChunks = (int)((h + lpc - 1) / lpc);
	std::vector<uint64_t> offsets(nChunks);
	for(int i = 0; i < nChunks; ++i)
		offsets[i] = r.u64();
	if(!r.ok) { err = "truncated offset table"; return false; }
	// pass 1 (serial): locate and validate every chunk; a table that was never written falls back to file order
	std::vector<ExrReader::Chunk>& refs = rd.chunks;
	refs.resize(nChunks);
	const uint8_t* seq = r.p;
	std::vector<bool> seen(nChunks, false); // every block of lines exactly once: no duplicates, no holes
	for(int i = 0; i < nChunks; ++i)
	{
		if(offsets[i] >= buf.size()) { err = "bad chunk offset"; return false; }
		const uint8_t* cp = offsets[i] ? buf.data() + offsets[i] : seq;
		Reader c(cp, (size_t)(buf.data() + buf.size() - cp));
		const int32_t y = c.i32(), n = c.i32();
		if(!c.ok || n < 0 || !c.need((size_t)n)) { err = "truncated chunk"; return false; }
		seq = c.p + n;
		const long ly = (long)y - out.minY;
		if(ly < 0 || ly >= h || ly % lpc) { err = "bad chunk coordinate"; return false; }
		if(seen[ly / lpc]) { err = "duplicate chunk"; return false; }
		seen[ly / lpc] = true;
		refs[i].offset = (size_t)(c.p - buf.data());
		refs[i].size = (size_t)n;
		refs[i].line = ly;
		refs[i].lines = (int)std::min<long>(lpc, h - ly);
		// what the chunk claims to expand to is bounded by what its bytes can encode (deflate tops out near 1032:1, the
		// run-length and Huffman coders below that), and an uncompressed chunk has exactly its size: a small file cannot
		// make the decoder — or the caller sizing its image from the header — allocate gigabytes
		const uint64_t rawSize = (uint64_t)lineBytes * refs[i].lines;
		if(out.compression == EXR_NONE ? rawSize != (uint64_t)n : rawSize > (uint64_t)n * 1100u + 4096u)
		{
			err = "chunk too small for the data window";
			return false;
		}
	}
	return true;
}

// Decompresses every chunk (in parallel: they are independent) and converts the channels to float. `planes` (one
// width x height plane per file channel) and / or `hwc` (interleaved, `depth` floats per pixel; slotOfChannel[c] is the
// position of file channel c inside a pixel or -1) receive the values.
static bool exrDecodeBody(const ExrReader& rd, std::vector<std::vector<float> >* planes, float* hwc, int depth,
		const std::vector<int>& slotOfChannel, std::string& err)
{
	const ExrFile& out = rd.info;
	const std::vector<ExrReader::Chunk>& refs = rd.chunks;
	const long w = out.width;
	const size_t nch = out.names.size(), lineBytes = rd.lineBytes;
	const int nChunks = (int)refs.size();
	bool failed = false;
	std::string firstErr;
#pragma omp parallel for schedule(dynamic, 1) if(nChunks > 4)
	for(int i = 0; i < nChunks; ++i)
	{
		if(failed)
			continue;
		std::string cerr;
		bool ok = true;
		try
		{
			const ExrReader::Chunk& cr = refs[i];
			const uint8_t* cdata = rd.file.data() + cr.offset;
			const size_t rawSize = lineBytes * cr.lines;
			std::vector<uint8_t> raw(rawSize);
			if(cr.size == rawSize)
				memcpy(raw.data(), cdata, rawSize);
			else
			{
				switch(out.compression)
				{
					case EXR_RLE: ok = unrleBlock(cdata, cr.size, rawSize, raw.data()); break;
					case EXR_ZIPS: case EXR_ZIP: ok = unzipBlock(cdata, cr.size, rawSize, raw.data()); break;
					case EXR_PIZ: ok = unpizBlock(cdata, cr.size, (int)w, cr.lines, out.types, raw.data(), cerr); break;
					default: ok = false; break;
				}
				if(!ok && cerr.empty()) cerr = "corrupt compressed chunk";
			}
			std::vector<float> row(hwc ? (size_t)w : 0);
			const uint8_t* s = raw.data();
			for(int l = 0; ok && l < cr.lines; ++l)
				for(size_t ch = 0; ch < nch; ++ch)
				{
					const int slot = hwc ? slotOfChannel[ch] : -1;
					float* dst = planes ? (*planes)[ch].data() + (size_t)(cr.line + l) * w : row.data();
					const size_t bytes = (size_t)w * bytesOf(out.types[ch]);
					if(!planes && slot < 0) { s += bytes; continue; } // channel nobody asked for
					if(out.types[ch] == EXR_HALF)
						for(long x = 0; x < w; ++x)
							dst[x] = halfToFloat((unsigned short)(s[2 * x] | s[2 * x + 1] << 8));
					else if(out.types[ch] == EXR_FLOAT)
						memcpy(dst, s, (size_t)w * 4); // little-endian file, little-endian host
					else
						for(long x = 0; x < w; ++x)
							dst[x] = (float)((uint32_t)s[4 * x] | (uint32_t)s[4 * x + 1] << 8 | (uint32_t)s[4 * x + 2] << 16 | (uint32_t)s[4 * x + 3] << 24);
					s += bytes;
					if(slot >= 0)
					{
						float* o = hwc + (size_t)(cr.line + l) * w * depth + slot;
						for(long x = 0; x < w; ++x)
							o[(size_t)x * depth] = dst[x];
					}
				}
		}
		catch(const std::bad_alloc&)
		{
			ok = false;
			cerr = "image too large for the available memory";
		}
		if(!ok)
		{
#pragma omp critical(bcdio_exr_read_error)
			{
				if(!failed) firstErr = cerr;
				failed = true;
			}
		}
	}
	if(failed)
	{
		err = firstErr;
		return false;
	}
	return true;
}

bool exrOpen(const char* path, ExrReader& rd, std::string& err)
{
	try
	{
		rd = ExrReader();
		return exrOpenBody(path, rd, err);
	}
	catch(const std::bad_alloc&)
	{
		rd = ExrReader();
		err = "image too large for the available memory";
		return false;
	}
}

bool exrDecodeInterleaved(const ExrReader& rd, const std::vector<int>& channelOfSlot, float* hwc, std::string& err)
{
	const int depth = (int)channelOfSlot.size();
	std::vector<int> slotOfChannel(rd.info.names.size(), -1);
	for(int z = 0; z < depth; ++z)
		if(channelOfSlot[z] >= 0 && channelOfSlot[z] < (int)slotOfChannel.size())
			slotOfChannel[channelOfSlot[z]] = z;
	return exrDecodeBody(rd, nullptr, hwc, depth, slotOfChannel, err);
}

bool exrRead(const char* path, ExrFile& out, std::string& err)
{
	try
	{
		ExrReader rd;
		if(!exrOpenBody(path, rd, err)) { out = ExrFile(); return false; }
		out = rd.info;
		out.planes.assign(out.names.size(), std::vector<float>((size_t)out.width * out.height));
		if(!exrDecodeBody(rd, &out.planes, nullptr, 0, std::vector<int>(), err)) { out = ExrFile(); return false; }
		return true;
	}
	catch(const std::bad_alloc&) // a header that announces more pixels than memory can hold
	{
		out = ExrFile();
		err = "image too large for the available memory";
		return false;
	}
}

// ------------------------------------------------------------------------------------------------------ write
namespace
{
void put32(std::vector<uint8_t>& b, uint32_t v)
{
	for(int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i)));
}
void putStr(std::vector<uint8_t>& b, const char* s)
{
	while(*s) b.push_back((uint8_t)*s++);
	b.push_back(0);
}
void putF32(std::vector<uint8_t>& b, float f)
{
	uint32_t v;
	memcpy(&v, &f, 4);
	put32(b, v);
}
void attr(std::vector<uint8_t>& b, const char* name, const char* type, const std::vector<uint8_t>& payload)
{
	putStr(b, name);
	putStr(b, type);
	put32(b, (uint32_t)payload.size());
	b.insert(b.end(), payload.begin(), payload.end());
}
} // namespace

bool exrWrite(const char* path, int width, int height, const std::vector<std::string>& names,
		const std::vector<const float*>& planes, ExrPixelType type, ExrCompression compression, std::string& err)
{
	if(width <= 0 || height <= 0 || names.empty() || names.size() != planes.size()) { err = "bad image for exrWrite"; return false; }
	if(type != EXR_HALF && type != EXR_FLOAT) { err = "exrWrite: HALF or FLOAT channels only"; return false; }
	if(compression != EXR_NONE && compression != EXR_ZIP && compression != EXR_ZIPS && compression != EXR_PIZ) { err = "exrWrite: NONE, ZIPS, ZIP or PIZ only"; return false; }
	std::vector<size_t> order(names.size()); // the channel list is alphabetical in the file
	for(size_t i = 0; i < order.size(); ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return names[a] < names[b]; });
	std::vector<uint8_t> head;
	put32(head, 20000630u);
	put32(head, 2u);
	{
		std::vector<uint8_t> p;
		for(size_t i : order)
		{
			putStr(p, names[i].c_str());
			put32(p, (uint32_t)type);
			put32(p, 0u); // pLinear + reserved
			put32(p, 1u);
			put32(p, 1u);
		}
		p.push_back(0);
		attr(head, "channels", "chlist", p);
	}
	attr(head, "compression", "compression", std::vector<uint8_t>(1, (uint8_t)compression));
	{
		std::vector<uint8_t> p;
		put32(p, 0u); put32(p, 0u); put32(p, (uint32_t)(width - 1)); put32(p, (uint32_t)(height - 1));
		attr(head, "dataWindow", "box2i", p);
		attr(head, "displayWindow", "box2i", p);
	}
	attr(head, "lineOrder", "lineOrder", std::vector<uint8_t>(1, 0));
	{
		std::vector<uint8_t> p;
		putF32(p, 1.f);
		attr(head, "pixelAspectRatio", "float", p);
		attr(head, "screenWindowWidth", "float", p);
		std::vector<uint8_t> c;
		putF32(c, 0.f); putF32(c, 0.f);
		attr(head, "screenWindowCenter", "v2f", c);
	}
	head.push_back(0);
	const int lpc = linesPerChunk(compression), nChunks = (height + lpc - 1) / lpc;
	const size_t px = type == EXR_HALF ? 2 : 4, lineBytes = (size_t)width * px * names.size();
	std::vector<std::vector<uint8_t> > chunks(nChunks);
	// the chunks are independent: pack and compress them on all host threads (deflate is the slow part)
#pragma omp parallel for schedule(dynamic, 1) if(nChunks > 4)
	for(int i = 0; i < nChunks; ++i)
	{
		std::vector<uint8_t> raw, t;
		const int y0 = i * lpc, lines = std::min(lpc, height - y0);
		raw.resize(lineBytes * lines);
		uint8_t* o = raw.data();
		for(int l = 0; l < lines; ++l)
			for(size_t ci : order)
			{
				const float* src = planes[ci] + (size_t)(y0 + l) * width;
				if(type == EXR_HALF)
					for(int x = 0; x < width; ++x)
					{
						const unsigned short hv = floatToHalf(src[x]);
						*o++ = (uint8_t)(hv & 0xff);
						*o++ = (uint8_t)(hv >> 8);
					}
				else
				{
					memcpy(o, src, (size_t)width * 4); // little-endian host
					o += (size_t)width * 4;
				}
			}
		std::vector<uint8_t>& c = chunks[i];
		put32(c, (uint32_t)y0);
		bool stored = false;
		if(compression == EXR_PIZ)
		{
			std::vector<uint8_t> z;
			pizBlock(raw.data(), width, lines, std::vector<int>(names.size(), (int)px), z);
			if(z.size() < raw.size())
			{
				put32(c, (uint32_t)z.size());
				c.insert(c.end(), z.begin(), z.end());
				stored = true;
			}
		}
		else if(compression != EXR_NONE)
		{
			applySplitAndPredictor(raw.data(), raw.size(), t);
			uLongf cap = compressBound((uLong)t.size());
			std::vector<uint8_t> z(cap);
			if(compress(z.data(), &cap, t.data(), (uLong)t.size()) == Z_OK && cap < raw.size())
			{
				put32(c, (uint32_t)cap);
				c.insert(c.end(), z.begin(), z.begin() + cap);
				stored = true;
			}
		}
		if(!stored)
		{
			put32(c, (uint32_t)raw.size());
			c.insert(c.end(), raw.begin(), raw.end());
		}
	}
	FILE* f = fopen(path, "wb");
	if(!f) { err = std::string("cannot write '") + path + "'"; return false; }
	bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
	uint64_t pos = head.size() + 8ull * nChunks;
	for(int i = 0; i < nChunks && ok; ++i)
	{
		uint8_t b[8];
		for(int k = 0; k < 8; ++k) b[k] = (uint8_t)(pos >> (8 * k));
		ok = fwrite(b, 1, 8, f) == 8;
		pos += chunks[i].size();
	}
	for(int i = 0; i < nChunks && ok; ++i)
		ok = fwrite(chunks[i].data(), 1, chunks[i].size(), f) == chunks[i].size();
	ok = (fclose(f) == 0) && ok;
	if(!ok) err = std::string("error while writing '") + path + "'";
	return ok;
}

} // namespace bcdio
