// tonemap.cuh — the consumer of the accumulation image: tone mapping + 8-bit encode.
//
// Reference: GltfRenderer::tonemap (src/renderer.cpp:992-1054) runs nvshaders::Tonemapper::runCompute on gBuffers[eImgRendered]
// (RGBA32F) into gBuffers[eImgTonemapped] (RGBA8), which saveHeadlessOutputImage writes to disk (src/renderer.cpp:557-573).  The
// compute shader and its TonemapperData live in nvpro_core2 (nvshaders/tonemap_*.slang), NOT in the reference tree: what is
// restated here are the PUBLISHED operators that shader offers, by their published formulas --
//   filmic      Hejl & Burgess-Dawson (2010), output already display-encoded
//   uncharted   Hable's Uncharted 2 curve, exposure bias 2, white point 11.2
//   clip        sRGB encode of the clamped colour
//   ACES        Stephen Hill's RRT+ODT fit
//   AgX         Wrensch's minimal AgX (Sobotka), default look
//   Khronos PBR neutral (Khronos 2024)
// followed by the classic post controls (contrast about 0.5, brightness as a gamma, saturation about Rec.601 luma, vignette) and
// the UNORM8 store.  Auto-exposure: the reference defaults to it (src/resources.hpp:212) through the external shader's
// luminance histogram; here it is the published log-average ("key value") form over a 256-bin log2-luminance histogram.
// Parity with the reference: UNPINNED (external source, no golden image in the tree); the oracle is oracle/tonemap.py.
#pragma once
#include "vec.cuh"

namespace pt {

enum : int
{
  TM_FILMIC = 0,
  TM_UNCHARTED = 1,
  TM_CLIP = 2,
  TM_ACES = 3,
  TM_AGX = 4,
  TM_KHRONOS_PBR = 5,
};

constexpr int   kTmBins = 256;
constexpr float kTmMinLog = -16.0f, kTmMaxLog = 16.0f;  // log2 luminance range of the histogram

PT_HD float tmSrgb1(float c)
{
  return c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
PT_HD float3 tmSrgb(float3 c) { return f3(tmSrgb1(fmaxf(c.x, 0.f)), tmSrgb1(fmaxf(c.y, 0.f)), tmSrgb1(fmaxf(c.z, 0.f))); }

PT_HD float tmFilmic1(float c)
{
  const float t = fmaxf(0.0f, c - 0.004f);
  return (t * (6.2f * t + 0.5f)) / (t * (6.2f * t + 1.7f) + 0.06f);
}
PT_HD float tmHable1(float x)
{
  const float a = 0.15f, b = 0.50f, c = 0.10f, d = 0.20f, e = 0.02f, f = 0.30f;
  return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f;
}
PT_HD float tmAgxContrast(float x)
{
  const float x2 = x * x, x4 = x2 * x2;
  return 15.5f * x4 * x2 - 40.14f * x4 * x + 31.96f * x4 - 6.868f * x2 * x + 0.4298f * x2 + 0.1191f * x - 0.00232f;
}

PT_HD float3 tonemapOperator(int method, float3 c)
{
  switch(method)
  {
    case TM_FILMIC:
      return f3(tmFilmic1(c.x), tmFilmic1(c.y), tmFilmic1(c.z));
    case TM_UNCHARTED: {
      const float ws = 1.0f / tmHable1(11.2f);
      return tmSrgb(f3(tmHable1(c.x * 2.0f) * ws, tmHable1(c.y * 2.0f) * ws, tmHable1(c.z * 2.0f) * ws));
    }
    case TM_ACES: {
      const float3 v = f3(0.59719f * c.x + 0.35458f * c.y + 0.04823f * c.z, 0.07600f * c.x + 0.90834f * c.y + 0.01566f * c.z,
                          0.02840f * c.x + 0.13383f * c.y + 0.83777f * c.z);
      auto         fit = [](float x) { return (x * (x + 0.0245786f) - 0.000090537f) / (x * (0.983729f * x + 0.4329510f) + 0.238081f); };
      const float3 w = f3(fit(v.x), fit(v.y), fit(v.z));
      return tmSrgb(f3(1.60475f * w.x - 0.53108f * w.y - 0.07367f * w.z, -0.10208f * w.x + 1.10813f * w.y - 0.00605f * w.z,
                       -0.00327f * w.x - 0.07276f * w.y + 1.07602f * w.z));
    }
    case TM_AGX: {
      const float  minEv = -12.47393f, maxEv = 4.026069f;
      const float3 v = f3(0.842479062253094f * c.x + 0.0784335999999992f * c.y + 0.0792237451477643f * c.z,
                          0.0423282422610123f * c.x + 0.878468636469772f * c.y + 0.0791661274605434f * c.z,
                          0.0423756549057051f * c.x + 0.0784336f * c.y + 0.879142973793104f * c.z);
      auto         enc = [&](float x) {
        const float l = fminf(fmaxf(log2f(fmaxf(x, 1e-10f)), minEv), maxEv);
        return tmAgxContrast((l - minEv) / (maxEv - minEv));
      };
      const float3 w = f3(enc(v.x), enc(v.y), enc(v.z));
      // inverse inset; the result is display-encoded (the minimal implementation's agxEotf without the final linearisation)
      return f3(1.19687900512017f * w.x - 0.0980208811401368f * w.y - 0.0990297440797205f * w.z,
                -0.0528968517574562f * w.x + 1.15190312990417f * w.y - 0.0989611768448433f * w.z,
                -0.0529716355144438f * w.x - 0.0980434501171241f * w.y + 1.15107367264116f * w.z);
    }
    case TM_KHRONOS_PBR: {
      const float startCompression = 0.8f - 0.04f, desaturation = 0.15f;
      const float x = fminf(c.x, fminf(c.y, c.z));
      const float offset = x < 0.08f ? x - 6.25f * x * x : 0.04f;
      float3      k = f3(c.x - offset, c.y - offset, c.z - offset);
      const float peak = fmaxf(k.x, fmaxf(k.y, k.z));
      if(peak < startCompression)
        return tmSrgb(k);
      const float d = 1.0f - startCompression;
      const float newPeak = 1.0f - d * d / (peak + d - startCompression);
      const float s = newPeak / peak;
      k = f3(k.x * s, k.y * s, k.z * s);
      const float g = 1.0f - 1.0f / (desaturation * (peak - newPeak) + 1.0f);
      return tmSrgb(f3(k.x + (newPeak - k.x) * g, k.y + (newPeak - k.y) * g, k.z + (newPeak - k.z) * g));
    }
    default:  // TM_CLIP
      return tmSrgb(c);
  }
}

// one pixel: exposure -> operator -> contrast / brightness / saturation / vignette; uv in [0,1]^2 (pixel centre / image size)
PT_HD float3 tonemapPixel(const b200pt_tonemapper& tm, float exposure, float3 c, float u, float v)
{
  if(!tm.isActive)
    return c;
  c = f3(c.x * exposure, c.y * exposure, c.z * exposure);
  float3 r = tonemapOperator(tm.method, c);
  auto   sat01 = [](float x) { return fminf(fmaxf(x, 0.0f), 1.0f); };
  r = f3(sat01(0.5f + (r.x - 0.5f) * tm.contrast), sat01(0.5f + (r.y - 0.5f) * tm.contrast), sat01(0.5f + (r.z - 0.5f) * tm.contrast));
  const float ib = 1.0f / tm.brightness;
  r = f3(powf(r.x, ib), powf(r.y, ib), powf(r.z, ib));
  const float luma = 0.299f * r.x + 0.587f * r.y + 0.114f * r.z;
  r = f3(luma + (r.x - luma) * tm.saturation, luma + (r.y - luma) * tm.saturation, luma + (r.z - luma) * tm.saturation);
  const float cu = (u - 0.5f) * 2.0f, cv = (v - 0.5f) * 2.0f;
  const float vg = 1.0f - (cu * cu + cv * cv) * tm.vignette;
  return f3(r.x * vg, r.y * vg, r.z * vg);
}

PT_HD uint32_t tmUnorm8(float x)
{
  x = fminf(fmaxf(x, 0.0f), 1.0f);  // (NaN -> 0 through fmaxf)
  return (uint32_t)(x * 255.0f + 0.5f);
}

// histogram bin of a luminance (bin 0 also takes everything at or below 2^kTmMinLog, black included)
PT_HD int tmBin(float lum)
{
  if(!(lum > 0.0f))
    return 0;
  const float t = (log2f(lum) - kTmMinLog) / (kTmMaxLog - kTmMinLog);
  const int   b = (int)(t * (float)kTmBins);
  return b < 0 ? 0 : (b >= kTmBins ? kTmBins - 1 : b);
}

}  // namespace pt
