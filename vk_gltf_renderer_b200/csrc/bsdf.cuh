// bsdf.cuh — device PBR BSDF evaluate / sample for the shade stage.
//
// Stands in for nvpro_core2/nvshaders/bsdf_functions.h.slang (EXTERNAL to the reference tree;
// call sites shaders/gltf_pathtrace.slang:333-350 `bsdfEvaluate`, :359-384 `bsdfSample`).
// Interface kept: k1 (to viewer), k2 (to light), xi (float3; xi.z picks ONE lobe stochastically),
// outputs bsdf_diffuse / bsdf_glossy (cosine included), pdf, bsdf_over_pdf, event_type.
// Lobes: diffuse reflection, diffuse transmission (KHR_materials_diffuse_transmission), rough
// dielectric transmission (KHR_materials_transmission/volume/ior), dielectric specular
// (KHR_materials_specular), metal, sheen (KHR_materials_sheen), clearcoat (KHR_materials_clearcoat);
// thin-film iridescence (KHR_materials_iridescence) tints the specular/metal lobes; anisotropic GGX
// through roughness.xy + T/B (KHR_materials_anisotropy).
#pragma once
#include "vec.cuh"

namespace pt {

struct PbrMaterial
{
  float3 baseColor;
  float  opacity;
  float2 roughness;  // GGX alpha (perceptual roughness squared)
  float  metallic;
  float3 emissive;
  float3 N, T, B, Ng;
  float  ior1, ior2;
  float  specular;
  float3 specularColor;
  float  transmission;
  float3 attenuationColor;
  float  attenuationDistance;
  float  thickness;
  float  clearcoat, clearcoatRoughness;
  float3 Nc;
  float  iridescence, iridescenceIor, iridescenceThickness;
  float3 sheenColor;
  float  sheenRoughness;
  float  diffuseTransmissionFactor;
  float3 diffuseTransmissionColor;
  float3 scatterCoefficient;
  float  scatterAnisotropy;
  float  dispersion;       // KHR_materials_dispersion (gltf_material_eval.h.slang:426-428)
  float  retroreflection;  // KHR_materials_retroreflection: carried like the reference does; b200pt_set_scene rejects factors > 0
};

// defaultPbrMaterial() (nvshaders pbr_material_types, external; restated like oracle/bsdf.h)
PT_D PbrMaterial defaultPbrMaterial()
{
  PbrMaterial m;
  m.baseColor = f3(1.0f);
  m.opacity = 1.0f;
  m.roughness = f2(1.0f, 1.0f);
  m.metallic = 1.0f;
  m.emissive = f3(0.0f);
  m.N = f3(0, 0, 1);
  m.T = f3(1, 0, 0);
  m.B = f3(0, 1, 0);
  m.Ng = f3(0, 0, 1);
  m.ior1 = 1.0f;
  m.ior2 = 1.5f;
  m.specular = 1.0f;
  m.specularColor = f3(1.0f);
  m.transmission = 0.0f;
  m.attenuationColor = f3(1.0f);
  m.attenuationDistance = 1.0f;
  m.thickness = 0.0f;
  m.clearcoat = 0.0f;
  m.clearcoatRoughness = 0.01f;
  m.Nc = f3(0, 0, 1);
  m.iridescence = 0.0f;
  m.iridescenceIor = 1.5f;
  m.iridescenceThickness = 0.1f;
  m.sheenColor = f3(0.0f);
  m.sheenRoughness = 0.0f;
  m.diffuseTransmissionFactor = 0.0f;
  m.diffuseTransmissionColor = f3(1.0f);
  m.scatterCoefficient = f3(0.0f);
  m.scatterAnisotropy = 0.0f;
  m.dispersion = 0.0f;
  m.retroreflection = 0.0f;
  return m;
}

enum : int
{
  BSDF_EVENT_ABSORB = 0,
  BSDF_EVENT_DIFFUSE = 1,
  BSDF_EVENT_GLOSSY = 1 << 1,
  BSDF_EVENT_IMPULSE = 1 << 2,
  BSDF_EVENT_REFLECTION = 1 << 3,
  BSDF_EVENT_TRANSMISSION = 1 << 4,
  BSDF_EVENT_DIFFUSE_REFLECTION = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION,
  BSDF_EVENT_DIFFUSE_TRANSMISSION = BSDF_EVENT_DIFFUSE | BSDF_EVENT_TRANSMISSION,
  BSDF_EVENT_GLOSSY_REFLECTION = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION,
  BSDF_EVENT_GLOSSY_TRANSMISSION = BSDF_EVENT_GLOSSY | BSDF_EVENT_TRANSMISSION,
};

// Per-scene kernel specialisation, like the reference's GLTF_USE_* shader variants
// (src/scene_shader_macros.cpp:40-56, shaders/gltf_eval_config.h:47-91): a feature the scene's materials never
// use is compiled out of the shade kernel.  Disabled features contribute exact zeros / identities, so the
// lean and the full variant produce bit-identical results on scenes that fit the lean one.
enum : uint32_t
{
  FEAT_TRANSMISSION = 1u << 0,
  FEAT_VOLUME = 1u << 1,
  FEAT_DIFFUSE_TRANSMISSION = 1u << 2,
  FEAT_CLEARCOAT = 1u << 3,
  FEAT_SHEEN = 1u << 4,
  FEAT_IRIDESCENCE = 1u << 5,
  FEAT_ANISOTROPY = 1u << 6,
  FEAT_SPECGLOSS = 1u << 7,
  FEAT_LIGHTS = 1u << 8,
  FEAT_ALL = 0x1ffu,
  // the variant Sponza / DamagedHelmet / Box class scenes run: metal-rough + specular + clearcoat + emissive + unlit
  FEAT_LEAN = FEAT_CLEARCOAT,
};

enum : int
{
  LOBE_DIFFUSE_REFLECTION = 0,
  LOBE_SPECULAR_TRANSMISSION = 1,
  LOBE_SPECULAR_REFLECTION = 2,
  LOBE_METAL_REFLECTION = 3,
  LOBE_SHEEN_REFLECTION = 4,
  LOBE_CLEARCOAT_REFLECTION = 5,
  LOBE_DIFFUSE_TRANSMISSION = 6,
  LOBE_COUNT = 7
};

PT_D float schlickFresnel(float ior, float cosTheta)
{
  float f0 = (ior - 1.0f) / (ior + 1.0f);
  f0 = f0 * f0;
  float m = 1.0f - cosTheta;
  float m2 = m * m;
  return f0 + (1.0f - f0) * (m2 * m2 * m);
}

PT_D float iorFresnel(float eta, float kh)
{
  float costheta = 1.0f - (1.0f - kh * kh) / (eta * eta);
  if(costheta <= 0.0f)
    return 1.0f;
  costheta = sqrtf(costheta);
  const float n2t1 = kh * eta;
  const float n2t2 = costheta * eta;
  const float r_p = (costheta - n2t1) / (costheta + n2t1);
  const float r_o = (kh - n2t2) / (kh + n2t2);
  return clampf(0.5f * (r_p * r_p + r_o * r_o), 0.0f, 1.0f);
}

PT_D float ggxD(float2 invRoughness, float3 h)
{
  const float x = h.x * invRoughness.x;
  const float y = h.y * invRoughness.y;
  const float f = (x * x + y * y) + h.z * h.z;
  return kInvPi * invRoughness.x * invRoughness.y * h.z / (f * f);
}

// Heitz 2018 visible-normal sampling
PT_D float3 ggxSampleVndf(float3 k, float2 roughness, float2 xi)
{
  const float3 v = normalize(f3(k.x * roughness.x, k.y * roughness.y, k.z));
  const float3 t1 = (v.z < 0.99999f) ? normalize(cross(v, f3(0, 0, 1))) : f3(1, 0, 0);
  const float3 t2 = cross(t1, v);
  const float  a = 1.0f / (1.0f + v.z);
  const float  r = sqrtf(xi.x);
  const float  phi = (xi.y < a) ? xi.y / a * kPi : kPi + (xi.y - a) / (1.0f - a) * kPi;
  const float  sp = sinf(phi);
  const float  cp = cosf(phi);
  const float  p1 = r * cp;
  const float  p2 = r * sp * ((xi.y < a) ? 1.0f : v.z);
  float3       h = t1 * p1 + t2 * p2 + v * sqrtf(fmaxf(0.0f, 1.0f - p1 * p1 - p2 * p2));
  h.x *= roughness.x;
  h.y *= roughness.y;
  h.z = fmaxf(0.0f, h.z);
  return normalize(h);
}

PT_D float smithG1(float3 k, float2 roughness)
{
  const float ax = k.x * roughness.x;
  const float ay = k.y * roughness.y;
  const float inv_a_2 = (ax * ax + ay * ay) / (k.z * k.z);
  return 2.0f / (1.0f + sqrtf(1.0f + inv_a_2));
}

__device__ __noinline__ float3 thinFilmFactor(float thickness, float coatIor, float baseIor, float inIor, float kh)
{
  const float cie[16][3] = {
      {0.02986f, 0.00310f, 0.13609f}, {0.20715f, 0.02304f, 0.99584f}, {0.36717f, 0.06469f, 1.89550f}, {0.28549f, 0.13661f, 1.67236f},
      {0.08233f, 0.26856f, 0.76653f}, {0.01723f, 0.48621f, 0.21889f}, {0.14400f, 0.77341f, 0.05886f}, {0.40957f, 0.95850f, 0.01280f},
      {0.74201f, 0.97967f, 0.00060f}, {1.03325f, 0.84591f, 0.00000f}, {1.08385f, 0.62242f, 0.00000f}, {0.79203f, 0.36749f, 0.00000f},
      {0.38751f, 0.16135f, 0.00000f}, {0.13401f, 0.05298f, 0.00000f}, {0.03531f, 0.01375f, 0.00000f}, {0.00817f, 0.00317f, 0.00000f}};
  thickness = fmaxf(0.0f, thickness);
  const float sin0_sqr = fmaxf(0.0f, 1.0f - kh * kh);
  const float eta01 = inIor / coatIor;
  const float sin1_sqr = eta01 * eta01 * sin0_sqr;
  if(sin1_sqr > 1.0f)
    return f3(1.0f);
  const float cos1 = sqrtf(fmaxf(0.0f, 1.0f - sin1_sqr));
  const float r01s = (inIor * kh - coatIor * cos1) / (inIor * kh + coatIor * cos1);
  const float r01p = (coatIor * kh - inIor * cos1) / (coatIor * kh + inIor * cos1);
  const float eta12 = coatIor / baseIor;
  const float sin2_sqr = eta12 * eta12 * sin1_sqr;
  float       r12s = 1.0f, r12p = 1.0f;
  if(sin2_sqr <= 1.0f)
  {
    const float cos2 = sqrtf(fmaxf(0.0f, 1.0f - sin2_sqr));
    r12s = (coatIor * cos1 - baseIor * cos2) / (coatIor * cos1 + baseIor * cos2);
    r12p = (baseIor * cos1 - coatIor * cos2) / (baseIor * cos1 + coatIor * cos2);
  }
  const float phaseK = 4.0f * kPi * coatIor * thickness * cos1;
  float       X = 0.0f, Y = 0.0f, Z = 0.0f, Xw = 0.0f, Yw = 0.0f, Zw = 0.0f;
  float       lambda = 400.0f;
#pragma unroll
  for(int i = 0; i < 16; ++i)
  {
    const float cphi = cosf(phaseK / lambda);
    const float ts = 2.0f * r01s * r12s * cphi;
    const float tp = 2.0f * r01p * r12p * cphi;
    const float Rs = (r01s * r01s + r12s * r12s + ts) / (1.0f + r01s * r01s * r12s * r12s + ts);
    const float Rp = (r01p * r01p + r12p * r12p + tp) / (1.0f + r01p * r01p * r12p * r12p + tp);
    const float R = 0.5f * (Rs + Rp);
    X += cie[i][0] * R;
    Y += cie[i][1] * R;
    Z += cie[i][2] * R;
    Xw += cie[i][0];
    Yw += cie[i][1];
    Zw += cie[i][2];
    lambda += 20.0f;
  }
  const float3 rgb = f3(3.2406f * X - 1.5372f * Y - 0.4986f * Z, -0.9689f * X + 1.8758f * Y + 0.0415f * Z, 0.0557f * X - 0.2040f * Y + 1.0570f * Z);
  const float3 white = f3(3.2406f * Xw - 1.5372f * Yw - 0.4986f * Zw, -0.9689f * Xw + 1.8758f * Yw + 0.0415f * Zw, 0.0557f * Xw - 0.2040f * Yw + 1.0570f * Zw);
  return f3(clampf(rgb.x / white.x, 0.0f, 1.0f), clampf(rgb.y / white.y, 0.0f, 1.0f), clampf(rgb.z / white.z, 0.0f, 1.0f));
}

PT_D float3 cosineSampleHemisphere(float r1, float r2)
{
  float  r = sqrtf(r1);
  float  phi = kTwoPi * r2;
  float3 dir;
  dir.x = r * cosf(phi);
  dir.y = r * sinf(phi);
  dir.z = sqrtf(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
  return dir;
}

PT_D float fresnelCosineApprox(float VdotN, float roughness) { return lerpf(VdotN, sqrtf(0.5f + 0.5f * VdotN), sqrtf(roughness)); }

// picks ONE lobe from the layered weights (clearcoat over sheen over metal | dielectric{spec, transmission, diffuse})
// lobeU: where inside the chosen lobe's probability interval rndVal fell, in [0, 1) -- a fresh uniform number that the
// transmission lobe uses to pick the colour channel of a dispersive refraction (only written for that lobe)
template <uint32_t FEAT>
PT_D int findLobe(const PbrMaterial& mat, float VdotN, float rndVal, float& lobeU)
{
  float frCoat = 0.0f;
  if((FEAT & FEAT_CLEARCOAT) && mat.clearcoat > 0.0f)
    frCoat = mat.clearcoat * iorFresnel(1.5f / mat.ior1, fresnelCosineApprox(VdotN, mat.clearcoatRoughness));
  float frDielectric = iorFresnel(mat.ior2 / mat.ior1, fresnelCosineApprox(VdotN, (mat.roughness.x + mat.roughness.y) * 0.5f));
  frDielectric *= mat.specular;
  float sheen = 0.0f;
  if((FEAT & FEAT_SHEEN) && (mat.sheenColor.x != 0.0f || mat.sheenColor.y != 0.0f || mat.sheenColor.z != 0.0f))
  {
    sheen = powf(1.0f - fabsf(VdotN), mat.sheenRoughness);
    sheen = sheen / (sheen + 0.5f);
  }
  const float base = (1.0f - frCoat) * (1.0f - sheen);
  const float diel = base * (1.0f - mat.metallic);
  const float diffuse = diel * (1.0f - frDielectric) * (1.0f - mat.transmission);
  // cumulative scan from the top lobe down (same order as the weights array walk); lobes of compiled-out
  // features have weight exactly 0 and are skipped
  float weight = 0.0f;
  if(FEAT & FEAT_DIFFUSE_TRANSMISSION)
  {
    weight = diffuse * mat.diffuseTransmissionFactor;  // LOBE_DIFFUSE_TRANSMISSION
    if(rndVal < weight)
      return LOBE_DIFFUSE_TRANSMISSION;
  }
  if(FEAT & FEAT_CLEARCOAT)
  {
    weight += frCoat;
    if(rndVal < weight)
      return LOBE_CLEARCOAT_REFLECTION;
  }
  if(FEAT & FEAT_SHEEN)
  {
    weight += (1.0f - frCoat) * sheen;
    if(rndVal < weight)
      return LOBE_SHEEN_REFLECTION;
  }
  weight += base * mat.metallic;
  if(rndVal < weight)
    return LOBE_METAL_REFLECTION;
  weight += diel * frDielectric;
  if(rndVal < weight)
    return LOBE_SPECULAR_REFLECTION;
  if(FEAT & FEAT_TRANSMISSION)
  {
    const float wT = diel * (1.0f - frDielectric) * mat.transmission;
    const float prev = weight;
    weight += wT;
    if(rndVal < weight)
    {
      lobeU = (rndVal - prev) / wT;
      return LOBE_SPECULAR_TRANSMISSION;
    }
  }
  return LOBE_DIFFUSE_REFLECTION;
}

// KHR_materials_dispersion on the specular-transmission lobe.  Per-channel index of refraction as the reference's own
// rasteriser defines it (shaders/gltf_raster.slang:204-208: halfSpread = (ior - 1) * 0.025 * dispersion, iors = {ior -
// halfSpread, ior, ior + halfSpread} for R, G, B); the path tracer follows ONE channel per refraction event, picked
// uniformly with lobeU, and weights it by 3 (an unbiased estimate of the three-channel sum).  The nvshaders body of
// this branch is not in the reference tree (unpinned like the rest of the BSDF stack).
PT_D float3 applyDispersion(PbrMaterial& mat, float lobeU)
{
  const int   c3 = (int)(lobeU * 3.0f);
  const int   channel = c3 < 2 ? c3 : 2;
  const float spread = (float)(channel - 1) * 0.025f * mat.dispersion;
  mat.ior1 = mat.ior1 + (mat.ior1 - 1.0f) * spread;  // the side that is air (ior 1) stays 1
  mat.ior2 = mat.ior2 + (mat.ior2 - 1.0f) * spread;
  return f3(channel == 0 ? 3.0f : 0.0f, channel == 1 ? 3.0f : 0.0f, channel == 2 ? 3.0f : 0.0f);
}

PT_D void iridescenceTint(const PbrMaterial& mat, int lobe, float kh, float3& tint)
{
  if(mat.iridescence > 0.0f)
  {
    const float3 factor = thinFilmFactor(mat.iridescenceThickness, mat.iridescenceIor, mat.ior2, mat.ior1, kh);
    if(lobe == LOBE_SPECULAR_REFLECTION)
      tint *= lerp3(f3(1.0f), factor, mat.iridescence);
    else if(lobe == LOBE_METAL_REFLECTION)
      tint = lerp3(tint, mat.specularColor * factor, mat.iridescence);
  }
}

// frame used by one microfacet lobe
struct LobeFrame
{
  float3 N, T, B;
  float2 roughness;
  float  iridescence;
};

PT_D LobeFrame lobeFrame(const PbrMaterial& mat, int lobe)
{
  LobeFrame f;
  if(lobe == LOBE_CLEARCOAT_REFLECTION)
  {
    const float a = mat.clearcoatRoughness * mat.clearcoatRoughness;
    f.roughness = f2(a, a);
    f.N = mat.Nc;
    f.B = normalize(cross(f.N, mat.T));
    f.T = cross(f.B, f.N);
    f.iridescence = 0.0f;
  }
  else
  {
    f.roughness = mat.roughness;
    f.N = mat.N;
    f.T = mat.T;
    f.B = mat.B;
    f.iridescence = mat.iridescence;
  }
  return f;
}

struct BsdfEval
{
  float3 bsdf_diffuse, bsdf_glossy;
  float  pdf;
};
struct BsdfSample
{
  float3 k2, bsdf_over_pdf;
  float  pdf;
  int    event_type;
};

template <uint32_t FEAT>
PT_D void ggxReflectEval(BsdfEval& d, const PbrMaterial& mat, const LobeFrame& fr, int lobe, float3 tint, float3 k1, float3 k2)
{
  if(dot(k2, mat.Ng) <= 0.0f)
    return;
  const float  nk1 = fabsf(dot(k1, fr.N));
  const float  nk2 = fabsf(dot(k2, fr.N));
  const float3 h = normalize(k1 + k2);
  const float  nh = dot(fr.N, h);
  const float  k1h = dot(k1, h);
  const float  k2h = dot(k2, h);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  const float3 h0 = f3(dot(fr.T, h), dot(fr.B, h), nh);
  float        pdf = ggxD(f2(1.0f / fr.roughness.x, 1.0f / fr.roughness.y), h0);
  const float  G1 = smithG1(f3(dot(fr.T, k1), dot(fr.B, k1), nk1), fr.roughness);
  const float  G2 = smithG1(f3(dot(fr.T, k2), dot(fr.B, k2), nk2), fr.roughness);
  pdf *= 0.25f / (nk1 * nh);
  const float3 bsdf = f3((G1 * G2) * pdf);
  d.pdf = pdf * G1;
  if((FEAT & FEAT_IRIDESCENCE) && fr.iridescence > 0.0f)
    iridescenceTint(mat, lobe, k1h, tint);
  d.bsdf_glossy = bsdf * tint;
}

template <uint32_t FEAT>
PT_D void ggxReflectSample(BsdfSample& d, const PbrMaterial& mat, const LobeFrame& fr, int lobe, float3 tint, float3 k1, float3 xi)
{
  const float nk1 = fabsf(dot(k1, fr.N));
  if(nk1 <= 0.0f)
    return;
  const float3 k10 = f3(dot(k1, fr.T), dot(k1, fr.B), nk1);
  const float3 h0 = ggxSampleVndf(k10, fr.roughness, f2(xi.x, xi.y));
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = fr.T * h0.x + fr.B * h0.y + fr.N * h0.z;
  const float  kh = dot(k1, h);
  if(kh <= 0.0f)
    return;
  const float3 k2 = h * (2.0f * kh) - k1;
  d.k2 = k2;
  if(dot(k2, mat.Ng) <= 0.0f)
    return;
  const float nk2 = fabsf(dot(k2, fr.N));
  const float G1 = smithG1(k10, fr.roughness);
  const float G2 = smithG1(f3(dot(k2, fr.T), dot(k2, fr.B), nk2), fr.roughness);
  const float G12 = G1 * G2;
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = f3(G12 / G1);
  d.pdf = ggxD(f2(1.0f / fr.roughness.x, 1.0f / fr.roughness.y), h0) * G1;
  d.pdf *= 0.25f / (nk1 * h0.z);
  if((FEAT & FEAT_IRIDESCENCE) && fr.iridescence > 0.0f)
    iridescenceTint(mat, lobe, kh, tint);
  d.bsdf_over_pdf *= tint;
  d.event_type = BSDF_EVENT_GLOSSY_REFLECTION;
}

PT_D void ggxTransmitEval(BsdfEval& d, const PbrMaterial& mat, float3 tint, float3 k1, float3 k2)
{
  const bool  thin = (mat.thickness == 0.0f);
  const float nk1 = fabsf(dot(k1, mat.N));
  const float nk2 = fabsf(dot(k2, mat.N));
  const bool  backside = (dot(k2, mat.Ng) <= 0.0f);
  float3      h;
  if(backside)
  {
    if(thin)
      h = k1 + (mat.N * (nk2 + nk2) + k2);
    else
    {
      h = k2 * mat.ior2 + k1 * mat.ior1;
      if(mat.ior2 > mat.ior1)
        h = h * -1.0f;
    }
  }
  else
    h = k1 + k2;
  h = normalize(h);
  const float nh = dot(mat.N, h);
  const float k1h = dot(k1, h);
  // thin-walled pseudo-BTDF: the half vector pairs k1 with the MIRRORED k2, so that direction's cosine is checked
  const float k2h = (backside && thin) ? dot(k2 + mat.N * (nk2 + nk2), h) : dot(k2, h) * (backside ? -1.0f : 1.0f);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  if(!backside)
  {
    const float b = mat.ior1 / mat.ior2;
    if(!(1.0f < (b * b * (1.0f - k1h * k1h))))
      return;  // only total internal reflection reflects in this lobe
  }
  const float3 h0 = f3(dot(mat.T, h), dot(mat.B, h), nh);
  float        pdf = ggxD(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0);
  const float  G1 = smithG1(f3(dot(mat.T, k1), dot(mat.B, k1), nk1), mat.roughness);
  const float  G2 = smithG1(f3(dot(mat.T, k2), dot(mat.B, k2), nk2), mat.roughness);
  if(!thin && backside)
  {
    const float tmp = k1h * mat.ior1 - k2h * mat.ior2;
    pdf *= k1h * k2h / (nk1 * nh * tmp * tmp);
  }
  else
    pdf *= 0.25f / (nk1 * nh);
  // prob == 1 on either branch that survives (fr = 1 for TIR reflection, 0 for transmission)
  const float3 bsdf = f3(1.0f * (G1 * G2) * pdf);
  d.pdf = pdf * (1.0f * G1);
  d.bsdf_glossy = bsdf * tint;
}

PT_D void ggxTransmitSample(BsdfSample& d, const PbrMaterial& mat, float3 tint, float3 k1, float3 xi)
{
  const bool   thin = (mat.thickness == 0.0f);
  const float  nk1 = fabsf(dot(k1, mat.N));
  const float3 k10 = f3(dot(k1, mat.T), dot(k1, mat.B), nk1);
  const float3 h0 = ggxSampleVndf(k10, mat.roughness, f2(xi.x, xi.y));
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
  const float  kh = dot(k1, h);
  if(kh <= 0.0f)
    return;
  bool   tir = false;
  float3 k2;
  if(thin)
  {
    float3 r = h * (2.0f * kh) - k1;
    k2 = normalize(r - mat.N * (2.0f * dot(r, mat.N)));
  }
  else
  {
    const float b = mat.ior1 / mat.ior2;
    const float refraction = b * b * (1.0f - kh * kh);
    tir = (1.0f <= refraction);
    k2 = tir ? (h * (kh + kh) - k1) : normalize(k1 * (-b) + h * (b * kh - sqrtf(1.0f - refraction)));
  }
  d.k2 = k2;
  const int   ev = tir ? BSDF_EVENT_GLOSSY_REFLECTION : BSDF_EVENT_GLOSSY_TRANSMISSION;
  const float gnk2 = dot(k2, mat.Ng) * (tir ? 1.0f : -1.0f);
  if(gnk2 <= 0.0f)
    return;
  const float nk2 = fabsf(dot(k2, mat.N));
  const float k2h = fabsf(dot(k2, h));
  const float G1 = smithG1(k10, mat.roughness);
  const float G2 = smithG1(f3(dot(k2, mat.T), dot(k2, mat.B), nk2), mat.roughness);
  const float G12 = G1 * G2;
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = f3(G12 / G1);
  d.pdf = ggxD(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0) * G1;
  if(!thin && !tir)
  {
    const float tmp = kh * mat.ior1 - k2h * mat.ior2;
    d.pdf *= kh * k2h / (nk1 * h0.z * tmp * tmp);
  }
  else
    d.pdf *= 0.25f / (nk1 * h0.z);
  d.bsdf_over_pdf *= tint;
  d.event_type = ev;
}

PT_D float sheenD(float invRoughness, float nh)
{
  const float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - nh * nh));
  return (invRoughness + 2.0f) * powf(sinTheta, invRoughness) * 0.5f * kInvPi * nh;
}
PT_D float vcavitiesMask(float nh, float kh, float nk) { return fminf(2.0f * nh * nk / kh, 1.0f); }

PT_D void sheenEval(BsdfEval& d, const PbrMaterial& mat, float3 k1, float3 k2)
{
  if(dot(k2, mat.Ng) <= 0.0f)
    return;
  const float  nk1 = fabsf(dot(k1, mat.N));
  const float  nk2 = fabsf(dot(k2, mat.N));
  const float3 h = normalize(k1 + k2);
  const float  nh = dot(mat.N, h);
  const float  k1h = dot(k1, h);
  const float  k2h = dot(k2, h);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  const float invRoughness = 1.0f / (mat.sheenRoughness * mat.sheenRoughness);
  float       pdf = sheenD(invRoughness, nh);
  const float G1 = vcavitiesMask(nh, k1h, nk1);
  const float G2 = vcavitiesMask(nh, k2h, nk2);
  pdf *= 0.25f / (nk1 * nh);
  const float3 bsdf = f3(pdf * fminf(G1, G2));
  d.pdf = pdf * G1;
  d.bsdf_glossy = bsdf * mat.sheenColor;
}

PT_D void sheenSample(BsdfSample& d, const PbrMaterial& mat, float3 k1, float3 xi)
{
  const float nk1 = fabsf(dot(k1, mat.N));
  if(nk1 <= 0.0f)
    return;
  const float3 k10 = f3(dot(k1, mat.T), dot(k1, mat.B), nk1);
  const float  invRoughness = 1.0f / (mat.sheenRoughness * mat.sheenRoughness);
  float3       h0;
  {
    const float phi = 2.0f * kPi * xi.x;
    const float sinPhi = sinf(phi);
    const float cosPhi = cosf(phi);
    const float sinTheta = powf(1.0f - xi.y, 1.0f / (invRoughness + 2.0f));
    const float cosTheta = sqrtf(fmaxf(0.0f, 1.0f - sinTheta * sinTheta));
    h0 = normalize(f3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
    // flip to the side of k1 with probability proportional to the projected area
    const float a = h0.z * k10.z;
    const float b = h0.x * k10.x + h0.y * k10.y;
    const float kh = fmaxf(0.0f, a + b);
    const float kh_f = fmaxf(0.0f, a - b);
    if(xi.z < kh_f / (kh + kh_f))
      h0 = f3(-h0.x, -h0.y, h0.z);
  }
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
  const float  k1h = dot(k1, h);
  if(k1h <= 0.0f)
    return;
  const float3 k2 = h * (2.0f * k1h) - k1;
  d.k2 = k2;
  if(dot(k2, mat.Ng) <= 0.0f)
    return;
  const float nk2 = fabsf(dot(k2, mat.N));
  const float k2h = fabsf(dot(k2, h));
  const float G1 = vcavitiesMask(h0.z, k1h, k10.z);
  const float G2 = vcavitiesMask(h0.z, k2h, nk2);
  const float G12 = fminf(G1, G2);
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = mat.sheenColor * (G12 / G1);
  d.pdf = sheenD(invRoughness, h0.z) * G1;
  d.pdf *= 0.25f / (nk1 * h0.z);
  d.event_type = BSDF_EVENT_GLOSSY_REFLECTION;
}

template <uint32_t FEAT>
__device__ __noinline__ BsdfEval bsdfEvaluate(const PbrMaterial& mat, float3 k1, float3 k2, float3 xi)
{
  BsdfEval d;
  d.bsdf_diffuse = f3(0.0f);
  d.bsdf_glossy = f3(0.0f);
  d.pdf = 0.0f;
  float     lobeU = 0.0f;
  const int lobe = findLobe<FEAT>(mat, dot(k1, mat.N), xi.z, lobeU);
  if(lobe == LOBE_DIFFUSE_REFLECTION)
  {
    if(dot(k2, mat.Ng) > 0.0f)
    {
      d.pdf = fmaxf(0.0f, dot(k2, mat.N) * kInvPi);
      d.bsdf_diffuse = mat.baseColor * d.pdf;
    }
  }
  else if((FEAT & FEAT_DIFFUSE_TRANSMISSION) && lobe == LOBE_DIFFUSE_TRANSMISSION)
  {
    if(dot(k2, mat.Ng) < 0.0f)
    {
      d.pdf = fmaxf(0.0f, -dot(k2, mat.N) * kInvPi);
      d.bsdf_diffuse = mat.diffuseTransmissionColor * d.pdf;
    }
  }
  else if((FEAT & FEAT_TRANSMISSION) && lobe == LOBE_SPECULAR_TRANSMISSION)
  {
    if(mat.dispersion > 0.0f)
    {
      PbrMaterial  md = mat;
      const float3 w = applyDispersion(md, lobeU);
      ggxTransmitEval(d, md, mat.baseColor * w, k1, k2);
    }
    else
      ggxTransmitEval(d, mat, mat.baseColor, k1, k2);
  }
  else if((FEAT & FEAT_SHEEN) && lobe == LOBE_SHEEN_REFLECTION)
    sheenEval(d, mat, k1, k2);
  else
  {
    const LobeFrame fr = lobeFrame(mat, lobe);
    const float3    tint = (lobe == LOBE_SPECULAR_REFLECTION) ? mat.specularColor : ((lobe == LOBE_METAL_REFLECTION) ? mat.baseColor : f3(1.0f));
    ggxReflectEval<FEAT>(d, mat, fr, lobe, tint, k1, k2);
  }
  return d;
}

template <uint32_t FEAT>
__device__ __noinline__ BsdfSample bsdfSample(const PbrMaterial& mat, float3 k1, float3 xi)
{
  BsdfSample d;
  d.k2 = f3(0.0f);
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  float     lobeU = 0.0f;
  const int lobe = findLobe<FEAT>(mat, dot(k1, mat.N), xi.z, lobeU);
  if(lobe == LOBE_DIFFUSE_REFLECTION || lobe == LOBE_DIFFUSE_TRANSMISSION)
  {
    const float  s = (lobe == LOBE_DIFFUSE_REFLECTION) ? 1.0f : -1.0f;
    const float3 l = cosineSampleHemisphere(xi.x, xi.y);
    d.k2 = normalize(mat.T * l.x + mat.B * l.y + mat.N * (s * l.z));
    d.pdf = s * dot(d.k2, mat.N) * kInvPi;
    d.bsdf_over_pdf = (lobe == LOBE_DIFFUSE_REFLECTION) ? mat.baseColor : mat.diffuseTransmissionColor;
    if(s > 0.0f)
      d.event_type = (0.0f < dot(d.k2, mat.Ng)) ? BSDF_EVENT_DIFFUSE_REFLECTION : BSDF_EVENT_ABSORB;
    else
      d.event_type = (dot(d.k2, mat.Ng) < 0.0f) ? BSDF_EVENT_DIFFUSE_TRANSMISSION : BSDF_EVENT_ABSORB;
  }
  else if((FEAT & FEAT_TRANSMISSION) && lobe == LOBE_SPECULAR_TRANSMISSION)
  {
    if(mat.dispersion > 0.0f)
    {
      PbrMaterial  md = mat;
      const float3 w = applyDispersion(md, lobeU);
      ggxTransmitSample(d, md, mat.baseColor * w, k1, xi);
    }
    else
      ggxTransmitSample(d, mat, mat.baseColor, k1, xi);
  }
  else if((FEAT & FEAT_SHEEN) && lobe == LOBE_SHEEN_REFLECTION)
    sheenSample(d, mat, k1, xi);
  else
  {
    const LobeFrame fr = lobeFrame(mat, lobe);
    const float3    tint = (lobe == LOBE_SPECULAR_REFLECTION) ? mat.specularColor : ((lobe == LOBE_METAL_REFLECTION) ? mat.baseColor : f3(1.0f));
    ggxReflectSample<FEAT>(d, mat, fr, lobe, tint, k1, xi);
  }
  if(d.pdf <= 0.00001f || isnan(d.bsdf_over_pdf.x) || isnan(d.bsdf_over_pdf.y) || isnan(d.bsdf_over_pdf.z))
    d.event_type = BSDF_EVENT_ABSORB;
  return d;
}

// ---- small nvshaders helpers used by hit fetch / volumes / environment --------------------------
// ---- bsdfEvaluateSimple / bsdfSampleSimple (nvshaders, external; call site: handleShadowCatcher, pathtrace_functions.h.slang:537-540) ----
// The reference continues a shadowed shadow-catcher path with this two-lobe model: Lambert diffuse plus ONE GGX lobe whose f0 blends
// the dielectric 0.04 and the base colour by metallic; the lobe is picked with the Schlick weight at N.V, the pair (value, pdf) then
// comes from evaluating BOTH lobes for the sampled direction.  Restated like the rest of this file (parity unpinned).
PT_D float schlickF0(float f0, float vdoth)
{
  const float m = 1.0f - vdoth, m2 = m * m;
  return f0 + (1.0f - f0) * (m2 * m2 * m);
}

PT_D BsdfEval bsdfEvaluateSimple(const PbrMaterial& mat, float3 k1, float3 k2)
{
  BsdfEval d;
  d.bsdf_diffuse = d.bsdf_glossy = f3(0.0f);
  d.pdf = 0.0f;
  const float3 h = normalize(k1 + k2);
  const float  nv = clampf(dot(mat.N, k1), 0.0f, 1.0f), nl = clampf(dot(mat.N, k2), 0.0f, 1.0f);
  const float  vh = clampf(dot(k1, h), 0.0f, 1.0f), nh = clampf(dot(mat.N, h), 0.0f, 1.0f);
  if(nv == 0.0f || nl == 0.0f || vh == 0.0f || nh == 0.0f)
    return d;
  const float  cMin = 0.04f;
  const float3 f0 = f3(cMin) + (mat.baseColor - f3(cMin)) * mat.metallic;
  const float3 fGlossy = f3(schlickF0(f0.x, vh), schlickF0(f0.y, vh), schlickF0(f0.z, vh));
  const float  fDiffuse = (1.0f - mat.metallic) * (1.0f - schlickF0(cMin, vh));
  const float3 h0 = f3(dot(mat.T, h), dot(mat.B, h), nh);
  const float  dd = ggxD(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0);
  const float  G1 = smithG1(f3(dot(mat.T, k1), dot(mat.B, k1), nv), mat.roughness);
  const float  G2 = smithG1(f3(dot(mat.T, k2), dot(mat.B, k2), nl), mat.roughness);
  const float  diffusePdf = kInvPi * nl;
  const float  specularPdf = G1 * dd * 0.25f / (nv * nh);
  d.pdf = specularPdf + (diffusePdf - specularPdf) * fDiffuse;
  d.bsdf_diffuse = mat.baseColor * (fDiffuse * diffusePdf);
  d.bsdf_glossy = fGlossy * (G2 * specularPdf);
  return d;
}

PT_D BsdfSample bsdfSampleSimple(const PbrMaterial& mat, float3 k1, float3 xi)
{
  BsdfSample d;
  d.k2 = f3(0.0f);
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  const float nv = clampf(dot(mat.N, k1), 0.0f, 1.0f);
  if(nv == 0.0f)
    return d;
  const float fDiffuse = (1.0f - mat.metallic) * (1.0f - schlickF0(0.04f, nv));
  int         ev;
  if(xi.z <= fDiffuse)
  {
    const float3 l = cosineSampleHemisphere(xi.x, xi.y);
    d.k2 = mat.T * l.x + mat.B * l.y + mat.N * l.z;
    ev = BSDF_EVENT_DIFFUSE_REFLECTION;
  }
  else
  {
    const float3 h0 = ggxSampleVndf(f3(dot(k1, mat.T), dot(k1, mat.B), nv), mat.roughness, f2(xi.x, xi.y));
    const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
    d.k2 = h * (2.0f * dot(k1, h)) - k1;
    ev = BSDF_EVENT_GLOSSY_REFLECTION;
  }
  if(dot(d.k2, mat.N) <= 0.0f)
    return d;
  const BsdfEval e = bsdfEvaluateSimple(mat, k1, d.k2);
  const float3   total = e.bsdf_diffuse + e.bsdf_glossy;
  if(!(e.pdf > 0.00001f) || total.x != total.x || total.y != total.y || total.z != total.z)
    return d;
  d.pdf = e.pdf;
  d.bsdf_over_pdf = total / e.pdf;
  d.event_type = ev;
  return d;
}

PT_D float3 pointOffset(float3 p, float3 pa, float3 pb, float3 pc, float3 na, float3 nb, float3 nc, float3 bary)
{
  float3      tu = p - pa, tv = p - pb, tw = p - pc;
  const float du = fminf(0.0f, dot(tu, na));
  const float dv = fminf(0.0f, dot(tv, nb));
  const float dw = fminf(0.0f, dot(tw, nc));
  tu -= na * du;
  tv -= nb * dv;
  tw -= nc * dw;
  return p + (tu * bary.x + tv * bary.y + tw * bary.z);
}

PT_D float4 makeFastTangent(float3 n)
{
  if(n.z < -0.99998796f)
    return f4(0.0f, -1.0f, 0.0f, 1.0f);
  const float a = 1.0f / (1.0f + n.z);
  const float b = -n.x * n.y * a;
  return f4(1.0f - n.x * n.x * a, b, -n.x, 1.0f);
}

PT_D float2 getSphericalUv(float3 v)
{
  const float gamma = asinf(-v.y);
  const float theta = atan2f(v.z, v.x);
  return f2(theta * (kInvPi * 0.5f) + 0.5f, gamma * kInvPi + 0.5f);
}

PT_D float3 rotateAxis(float3 v, float3 k, float theta)
{
  const float c = cosf(theta), s = sinf(theta);
  return (v * c) + (cross(k, v) * s) + (k * dot(k, v)) * (1.0f - c);
}

PT_D float henyeyGreensteinPdf(float cosTheta, float g)
{
  const float denom = 1.0f + g * g - 2.0f * g * cosTheta;
  return (1.0f / (4.0f * kPi)) * (1.0f - g * g) / (denom * sqrtf(denom));
}
PT_D float3 sampleHenyeyGreenstein(float2 xi, float g, float3 wi)
{
  float cosTheta;
  if(fabsf(g) < 1e-3f)
    cosTheta = 1.0f - 2.0f * xi.x;
  else
  {
    const float sq = (1.0f - g * g) / (1.0f - g + 2.0f * g * xi.x);
    cosTheta = (1.0f + g * g - sq * sq) / (2.0f * g);
  }
  cosTheta = clampf(cosTheta, -1.0f, 1.0f);
  const float  sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
  const float  phi = kTwoPi * xi.y;
  const float3 T = normalize(xyz(makeFastTangent(wi)));
  const float3 B = cross(wi, T);
  return normalize(T * (sinTheta * cosf(phi)) + B * (sinTheta * sinf(phi)) + wi * cosTheta);
}

}  // namespace pt
