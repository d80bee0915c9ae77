// Robustness harness for the from-scratch EXR reader (bcd_b200/io/exr_codec.cpp): every file in argv[3..] is mutated
// argv[1] times (byte flips, truncation, overwritten length / coordinate fields, spliced chunks) with the seed argv[2] and
// fed to exrRead, which must return true or false — never crash, hang, or trip AddressSanitizer / UBSan (the harness is
// compiled with -fsanitize=address,undefined by tests/test_exr_io.py).
#include "exr_codec.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static uint64_t g_state;
static uint32_t rnd()
{
	g_state = g_state * 6364136223846793005ull + 1442695040888963407ull;
	return (uint32_t)(g_state >> 33);
}

int main(int argc, char** argv)
{
	if(argc < 4) return 2;
	const int iterations = atoi(argv[1]);
	g_state = strtoull(argv[2], nullptr, 10) * 2654435761ull + 12345;
	const std::string scratch = std::string(argv[3]);
	int accepted = 0, rejected = 0;
	for(int f = 4; f < argc; ++f)
	{
		std::vector<uint8_t> good;
		FILE* in = fopen(argv[f], "rb");
		if(!in) return 3;
		int c;
		while((c = fgetc(in)) != EOF) good.push_back((uint8_t)c);
		fclose(in);
		for(int it = 0; it < iterations; ++it)
		{
			std::vector<uint8_t> b(good);
			const uint32_t kind = rnd() % 6;
			if(kind == 0) // a few random byte flips anywhere
				for(uint32_t k = 0, n = 1 + rnd() % 8; k < n; ++k) b[rnd() % b.size()] ^= (uint8_t)(1u << (rnd() % 8));
			else if(kind == 1) // truncation
				b.resize(rnd() % b.size());
			else if(kind == 2) // garbage in the header
				for(uint32_t k = 0, n = 1 + rnd() % 16; k < n; ++k) b[rnd() % (b.size() < 400 ? b.size() : 400)] = (uint8_t)rnd();
			else if(kind == 3) // a 32-bit field overwritten with an extreme value
			{
				static const uint32_t extreme[] = {0u, 1u, 0x7fffffffu, 0x80000000u, 0xffffffffu, 0x00010000u, 0xfffffff0u};
				const size_t at = rnd() % (b.size() - 4);
				const uint32_t v = extreme[rnd() % 7];
				for(int k = 0; k < 4; ++k) b[at + k] = (uint8_t)(v >> (8 * k));
			}
			else if(kind == 4) // a run of random bytes inside the pixel data
			{
				const size_t at = b.size() / 3 + rnd() % (b.size() / 2);
				for(size_t k = at; k < b.size() && k < at + 64; ++k) b[k] = (uint8_t)rnd();
			}
			else // a block copied over another place
			{
				const size_t n = 1 + rnd() % 200, from = rnd() % (b.size() - n), to = rnd() % (b.size() - n);
				for(size_t k = 0; k < n; ++k) b[to + k] = b[from + k];
			}
			FILE* out = fopen(scratch.c_str(), "wb");
			if(!out) return 4;
			fwrite(b.data(), 1, b.size(), out);
			fclose(out);
			bcdio::ExrFile img;
			std::string err;
			if(bcdio::exrRead(scratch.c_str(), img, err)) ++accepted; else ++rejected;
		}
	}
	printf("accepted %d rejected %d\n", accepted, rejected);
	return 0;
}
