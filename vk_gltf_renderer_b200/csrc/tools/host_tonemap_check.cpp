// host_tonemap_check.cpp — the device tone-mapping source (tonemap.cuh: tonemapPixel / tmUnorm8 / tmBin, the bodies of k_tonemap
// and k_tm_histogram) compiled for the host and run pixel by pixel on an image dumped by tests/test_tonemap.py.
//   host_tonemap_check in.bin out.bin
// in.bin: u32 width, rows, y0, fullHeight; b200pt_tonemapper (32 bytes); f32 exposure; RGBA32F pixels.
// out.bin: RGBA8 pixels, then 256 x u32 histogram.
#include "host_shim.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../../include/b200pt.h"
#include "../tonemap.cuh"

using namespace pt;

int main(int argc, char** argv)
{
  if(argc < 3)
    return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if(!f)
    return 2;
  uint32_t          hd[4];
  b200pt_tonemapper tm;
  float             exposure;
  if(std::fread(hd, 4, 4, f) != 4 || std::fread(&tm, sizeof(tm), 1, f) != 1 || std::fread(&exposure, 4, 1, f) != 1)
    return 2;
  const uint32_t     n = hd[0] * hd[1];
  std::vector<float> img((size_t)n * 4);
  if(std::fread(img.data(), 4, img.size(), f) != img.size())
    return 2;
  std::fclose(f);
  std::vector<uint8_t>  out((size_t)n * 4);
  std::vector<uint32_t> hist(kTmBins, 0u);
  for(uint32_t i = 0; i < n; i++)
  {
    const float* c = &img[(size_t)i * 4];
    const int    x = (int)(i % hd[0]), y = (int)hd[2] + (int)(i / hd[0]);
    const float3 r = tonemapPixel(tm, exposure, f3(c[0], c[1], c[2]), ((float)x + 0.5f) / (float)hd[0], ((float)y + 0.5f) / (float)hd[3]);
    out[(size_t)i * 4 + 0] = (uint8_t)tmUnorm8(r.x);
    out[(size_t)i * 4 + 1] = (uint8_t)tmUnorm8(r.y);
    out[(size_t)i * 4 + 2] = (uint8_t)tmUnorm8(r.z);
    out[(size_t)i * 4 + 3] = (uint8_t)tmUnorm8(c[3]);
    hist[tmBin(0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2])]++;
  }
  FILE* o = std::fopen(argv[2], "wb");
  if(!o)
    return 2;
  std::fwrite(out.data(), 1, out.size(), o);
  std::fwrite(hist.data(), 4, hist.size(), o);
  std::fclose(o);
  return 0;
}
