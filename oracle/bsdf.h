// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/pt_oracle.cpp header).
//
// Restatement of the EXTERNAL arithmetic the reference path tracer calls:
//   nvpro_core2/nvshaders/{bsdf_types,bsdf_functions,pbr_material_types,pbr_ggx_microfacet,
//   random,functions,ray_utils,light_contrib}.h.slang   — NOT under /root/reference
//   (fetched at configure time from branch `main`, unpinned: cmake/FindNvproCore2.cmake:28,85).
//
// PARITY UNPINNED: the bodies below restate the published algorithms those headers implement
// (MDL-SDK libbsdf microfacet/sheen/thin-film models as used by nvpro_core; Heitz 2018 VNDF
// sampling; Jarzynski-Olano xxhash32 + PCG; Hanika 2021 shadow-terminator offset; Vose alias
// tables) from their documented interfaces at the reference's call sites:
//   bsdfEvaluate  shaders/gltf_pathtrace.slang:333-350      bsdfSample  :359-384
//   xxhash32/rand :560,336,361                              pointOffset get_hit.h.slang:105
//   makeFastTangent get_hit.h.slang:139                     schlickFresnel pathtrace_functions.h.slang:283
//   HG            pathtrace_functions.h.slang:625-627,660   singleLightContribution :406
// No golden vector of the reference exists for any of them.
#pragma once
#include "vecmath.h"

namespace orc {

static const float M_PI_F = 3.14159265358979323846f;
static const float M_TWO_PI_F = 6.28318530717958647692f;
static const float M_1_PI_F = 0.31830988618379067154f;
static const float INFINITE_F = 1e32f;  // nvshaders/constants: payload.hitT == INFINITE (gltf_pathtrace.slang:112)
static const float DIRAC = -1.0f;       // pdf sentinel for delta events (gltf_pathtrace.slang:344)

// ---------------------------------------------------------------------------------------------
// RNG (nvshaders/random.h.slang)
// ---------------------------------------------------------------------------------------------
static inline uint32_t xxhash32(uint32_t px, uint32_t py, uint32_t pz)
{
  const uint32_t P1 = 2246822519u, P2 = 3266489917u, P3 = 668265263u, P4 = 374761393u;
  uint32_t h32 = pz + P4 + px * P2;
  h32 = P3 * ((h32 << 17) | (h32 >> 15));
  h32 += py * P2;
  h32 = P3 * ((h32 << 17) | (h32 >> 15));
  h32 = P1 * (h32 ^ (h32 >> 15));
  h32 = P2 * (h32 ^ (h32 >> 13));
  return h32 ^ (h32 >> 16);
}
static inline uint32_t pcg(uint32_t& state)
{
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
static inline float rnd(uint32_t& seed)
{
  uint32_t r = pcg(seed);
  return asfloat(0x3f800000u | (r >> 9)) - 1.0f;
}

// ---------------------------------------------------------------------------------------------
// PbrMaterial (nvshaders/pbr_material_types.h.slang) — fields the reference reads/writes
// (gltf_material_eval.h.slang:172-456)
// ---------------------------------------------------------------------------------------------
struct PbrMaterial
{
  float3 baseColor;
  float  opacity;
  float2 roughness;  // alpha (already squared)
  float  metallic;
  float3 emissive;
  float  occlusion;
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
  float  dispersion;
  float  diffuseTransmissionFactor;
  float3 diffuseTransmissionColor;
  float3 scatterCoefficient;
  float  scatterAnisotropy;
  float  retroreflection;
};

static inline PbrMaterial defaultPbrMaterial()
{
  PbrMaterial m;
  m.baseColor = f3(1.0f);
  m.opacity = 1.0f;
  m.roughness = f2(1.0f, 1.0f);
  m.metallic = 1.0f;
  m.emissive = f3(0.0f);
  m.occlusion = 1.0f;
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
  m.dispersion = 0.0f;
  m.diffuseTransmissionFactor = 0.0f;
  m.diffuseTransmissionColor = f3(1.0f);
  m.scatterCoefficient = f3(0.0f);
  m.scatterAnisotropy = 0.0f;
  m.retroreflection = 0.0f;
  return m;
}

// ---------------------------------------------------------------------------------------------
// BSDF data + events (nvshaders/bsdf_types.h.slang)
// ---------------------------------------------------------------------------------------------
enum
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
  BSDF_EVENT_IMPULSE_REFLECTION = BSDF_EVENT_IMPULSE | BSDF_EVENT_REFLECTION,
  BSDF_EVENT_IMPULSE_TRANSMISSION = BSDF_EVENT_IMPULSE | BSDF_EVENT_TRANSMISSION,
};

struct BsdfEvaluateData
{
  float3 k1, k2, xi;
  float3 bsdf_diffuse, bsdf_glossy;
  float  pdf;
};
struct BsdfSampleData
{
  float3 k1, k2, xi;
  float  pdf;
  float3 bsdf_over_pdf;
  int    event_type;
};

enum
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

// ---- microfacet helpers (MDL libbsdf as used by nvshaders/pbr_ggx_microfacet) -----------------
static inline float schlickFresnel(float ior, float cosTheta)
{
  float f0 = (ior - 1.0f) / (ior + 1.0f);
  f0 = f0 * f0;
  float m = 1.0f - cosTheta;
  float m2 = m * m;
  return f0 + (1.0f - f0) * (m2 * m2 * m);
}

// Fresnel for an equal mix of polarisations; eta = n_transmitted / n_incident
static inline float ior_fresnel(float eta, float kh)
{
  float costheta = 1.0f - (1.0f - kh * kh) / (eta * eta);
  if(costheta <= 0.0f)
    return 1.0f;
  costheta = sqrtf(costheta);
  const float n1t1 = kh;
  const float n1t2 = costheta;
  const float n2t1 = kh * eta;
  const float n2t2 = costheta * eta;
  const float r_p = (n1t2 - n2t1) / (n1t2 + n2t1);
  const float r_o = (n1t1 - n2t2) / (n1t1 + n2t2);
  const float fres = 0.5f * (r_p * r_p + r_o * r_o);
  return clampf(fres, 0.0f, 1.0f);
}

static inline bool isTIR(float ior1, float ior2, float kh)
{
  const float b = ior1 / ior2;
  return 1.0f < (b * b * (1.0f - kh * kh));
}

static inline float hvd_ggx_eval(float2 invRoughness, float3 h)
{
  const float x = h.x * invRoughness.x;
  const float y = h.y * invRoughness.y;
  const float aniso = x * x + y * y;
  const float f = aniso + h.z * h.z;
  return M_1_PI_F * invRoughness.x * invRoughness.y * h.z / (f * f);
}

// Heitz 2018, "Sampling the GGX distribution of visible normals"
static inline float3 hvd_ggx_sample_vndf(float3 k, float2 roughness, float2 xi)
{
  const float3 v = normalize(f3(k.x * roughness.x, k.y * roughness.y, k.z));
  const float3 t1 = (v.z < 0.99999f) ? normalize(cross(v, f3(0, 0, 1))) : f3(1, 0, 0);
  const float3 t2 = cross(t1, v);
  const float  a = 1.0f / (1.0f + v.z);
  const float  r = sqrtf(xi.x);
  const float  phi = (xi.y < a) ? xi.y / a * M_PI_F : M_PI_F + (xi.y - a) / (1.0f - a) * M_PI_F;
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

static inline float smith_shadow_mask(float3 k, float2 roughness)
{
  const float ax = k.x * roughness.x;
  const float ay = k.y * roughness.y;
  const float inv_a_2 = (ax * ax + ay * ay) / (k.z * k.z);
  return 2.0f / (1.0f + sqrtf(1.0f + inv_a_2));
}
static inline float ggx_smith_shadow_mask(float& G1, float& G2, float3 k1, float3 k2, float2 roughness)
{
  G1 = smith_shadow_mask(k1, roughness);
  G2 = smith_shadow_mask(k2, roughness);
  return G1 * G2;
}

static inline float3 refract_h(float3 k, float3 n, float b, float nk, bool& tir)
{
  const float refraction = b * b * (1.0f - nk * nk);
  tir = (1.0f <= refraction);
  return tir ? (n * (nk + nk) - k) : normalize(k * (-b) + n * (b * nk - sqrtf(1.0f - refraction)));
}

static inline float3 compute_half_vector(float3 k1, float3 k2, float3 normal, float ior1, float ior2, float nk2, bool transmission, bool thinwalled)
{
  float3 h;
  if(transmission)
  {
    if(thinwalled)
      h = k1 + (normal * (nk2 + nk2) + k2);
    else
    {
      h = k2 * ior2 + k1 * ior1;
      if(ior2 > ior1)
        h = h * -1.0f;
    }
  }
  else
    h = k1 + k2;
  return normalize(h);
}

// sheen (MDL sheen_bsdf)
static inline float hvd_sheen_eval(float invRoughness, float nh)
{
  const float sinTheta2 = fmaxf(0.0f, 1.0f - nh * nh);
  const float sinTheta = sqrtf(sinTheta2);
  return (invRoughness + 2.0f) * powf(sinTheta, invRoughness) * 0.5f * M_1_PI_F * nh;
}
static inline float vcavities_mask(float nh, float kh, float nk) { return fminf(2.0f * nh * nk / kh, 1.0f); }
static inline float vcavities_shadow_mask(float& G1, float& G2, float nh, float3 k1, float k1h, float3 k2, float k2h)
{
  G1 = vcavities_mask(nh, k1h, k1.z);
  G2 = vcavities_mask(nh, k2h, k2.z);
  return fminf(G1, G2);
}
static inline float3 hvd_sheen_sample(float2 xi, float invRoughness)
{
  const float phi = 2.0f * M_PI_F * xi.x;
  const float sinPhi = sinf(phi);
  const float cosPhi = cosf(phi);
  const float sinTheta = powf(1.0f - xi.y, 1.0f / (invRoughness + 2.0f));
  const float cosTheta = sqrtf(fmaxf(0.0f, 1.0f - sinTheta * sinTheta));
  return normalize(f3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
}
static inline float3 flip(float3 h, float3 k, float xi)
{
  const float a = h.z * k.z;
  const float b = h.x * k.x + h.y * k.y;
  const float kh = fmaxf(0.0f, a + b);
  const float kh_f = fmaxf(0.0f, a - b);
  const float p_flip = kh_f / (kh + kh_f);
  if(xi < p_flip)
    return f3(-h.x, -h.y, h.z);
  return h;
}

// thin-film interference (MDL libbsdf thin_film_factor): Airy reflectance of a single coating
// layer averaged over 16 wavelengths, projected through CIE XYZ to linear sRGB, normalised so a
// perfect mirror (R == 1) maps to (1,1,1).
static inline float3 thin_film_factor(float coating_thickness, float coating_ior, float base_ior, float incoming_ior, float kh)
{
  static const float cie_xyz[16][3] = {
      {0.02986f, 0.00310f, 0.13609f}, {0.20715f, 0.02304f, 0.99584f}, {0.36717f, 0.06469f, 1.89550f}, {0.28549f, 0.13661f, 1.67236f},
      {0.08233f, 0.26856f, 0.76653f}, {0.01723f, 0.48621f, 0.21889f}, {0.14400f, 0.77341f, 0.05886f}, {0.40957f, 0.95850f, 0.01280f},
      {0.74201f, 0.97967f, 0.00060f}, {1.03325f, 0.84591f, 0.00000f}, {1.08385f, 0.62242f, 0.00000f}, {0.79203f, 0.36749f, 0.00000f},
      {0.38751f, 0.16135f, 0.00000f}, {0.13401f, 0.05298f, 0.00000f}, {0.03531f, 0.01375f, 0.00000f}, {0.00817f, 0.00317f, 0.00000f}};
  coating_thickness = fmaxf(0.0f, coating_thickness);
  const float sin0_sqr = fmaxf(0.0f, 1.0f - kh * kh);
  const float eta01 = incoming_ior / coating_ior;
  const float sin1_sqr = eta01 * eta01 * sin0_sqr;
  if(sin1_sqr > 1.0f)
    return f3(1.0f);  // TIR at the first interface
  const float cos1 = sqrtf(fmaxf(0.0f, 1.0f - sin1_sqr));
  // amplitude coefficients, interface 0 -> 1
  const float r01s = (incoming_ior * kh - coating_ior * cos1) / (incoming_ior * kh + coating_ior * cos1);
  const float r01p = (coating_ior * kh - incoming_ior * cos1) / (coating_ior * kh + incoming_ior * cos1);
  // interface 1 -> 2
  const float eta12 = coating_ior / base_ior;
  const float sin2_sqr = eta12 * eta12 * sin1_sqr;
  float       r12s = 1.0f, r12p = 1.0f;
  if(sin2_sqr <= 1.0f)
  {
    const float cos2 = sqrtf(fmaxf(0.0f, 1.0f - sin2_sqr));
    r12s = (coating_ior * cos1 - base_ior * cos2) / (coating_ior * cos1 + base_ior * cos2);
    r12p = (base_ior * cos1 - coating_ior * cos2) / (base_ior * cos1 + coating_ior * cos2);
  }
  const float phase_k = 4.0f * M_PI_F * coating_ior * coating_thickness * cos1;
  float       X = 0.0f, Y = 0.0f, Z = 0.0f, Xw = 0.0f, Yw = 0.0f, Zw = 0.0f;
  float       lambda = 400.0f;
  for(int i = 0; i < 16; ++i)
  {
    const float cphi = cosf(phase_k / lambda);
    const float ts = 2.0f * r01s * r12s * cphi;
    const float tp = 2.0f * r01p * r12p * cphi;
    const float Rs = (r01s * r01s + r12s * r12s + ts) / (1.0f + r01s * r01s * r12s * r12s + ts);
    const float Rp = (r01p * r01p + r12p * r12p + tp) / (1.0f + r01p * r01p * r12p * r12p + tp);
    const float R = 0.5f * (Rs + Rp);
    X += cie_xyz[i][0] * R;
    Y += cie_xyz[i][1] * R;
    Z += cie_xyz[i][2] * R;
    Xw += cie_xyz[i][0];
    Yw += cie_xyz[i][1];
    Zw += cie_xyz[i][2];
    lambda += 20.0f;
  }
  const float3 rgb = f3(3.2406f * X - 1.5372f * Y - 0.4986f * Z, -0.9689f * X + 1.8758f * Y + 0.0415f * Z, 0.0557f * X - 0.2040f * Y + 1.0570f * Z);
  const float3 white = f3(3.2406f * Xw - 1.5372f * Yw - 0.4986f * Zw, -0.9689f * Xw + 1.8758f * Yw + 0.0415f * Zw, 0.0557f * Xw - 0.2040f * Yw + 1.0570f * Zw);
  return f3(clampf(rgb.x / white.x, 0.0f, 1.0f), clampf(rgb.y / white.y, 0.0f, 1.0f), clampf(rgb.z / white.z, 0.0f, 1.0f));
}

static inline float3 cosineSampleHemisphere(float r1, float r2)
{
  float  r = sqrtf(r1);
  float  phi = M_TWO_PI_F * r2;
  float3 dir;
  dir.x = r * cosf(phi);
  dir.y = r * sinf(phi);
  dir.z = sqrtf(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
  return dir;
}

// ---- lobe selection ---------------------------------------------------------------------------
static inline float fresnelCosineApproximation(float VdotN, float roughness)
{
  return lerpf(VdotN, sqrtf(0.5f + 0.5f * VdotN), sqrtf(roughness));
}

static inline void computeLobeWeights(const PbrMaterial& mat, float VdotN, float w[LOBE_COUNT])
{
  float frCoat = 0.0f;
  if(mat.clearcoat > 0.0f)
  {
    float frCosineClearcoat = fresnelCosineApproximation(VdotN, mat.clearcoatRoughness);
    frCoat = mat.clearcoat * ior_fresnel(1.5f / mat.ior1, frCosineClearcoat);
  }
  float frCosine = fresnelCosineApproximation(VdotN, (mat.roughness.x + mat.roughness.y) * 0.5f);
  float frDielectric = ior_fresnel(mat.ior2 / mat.ior1, frCosine);
  frDielectric *= mat.specular;

  float sheen = 0.0f;
  if(mat.sheenColor.x != 0.0f || mat.sheenColor.y != 0.0f || mat.sheenColor.z != 0.0f)
  {
    sheen = powf(1.0f - fabsf(VdotN), mat.sheenRoughness);
    sheen = sheen / (sheen + 0.5f);
  }
  const float base = (1.0f - frCoat) * (1.0f - sheen);
  const float diel = base * (1.0f - mat.metallic);
  const float diffuse = diel * (1.0f - frDielectric) * (1.0f - mat.transmission);
  w[LOBE_CLEARCOAT_REFLECTION] = frCoat;
  w[LOBE_SHEEN_REFLECTION] = (1.0f - frCoat) * sheen;
  w[LOBE_METAL_REFLECTION] = base * mat.metallic;
  w[LOBE_SPECULAR_REFLECTION] = diel * frDielectric;
  w[LOBE_SPECULAR_TRANSMISSION] = diel * (1.0f - frDielectric) * mat.transmission;
  w[LOBE_DIFFUSE_TRANSMISSION] = diffuse * mat.diffuseTransmissionFactor;
  w[LOBE_DIFFUSE_REFLECTION] = diffuse * (1.0f - mat.diffuseTransmissionFactor);
}

// lobeU: position of rndVal inside the chosen lobe's probability interval, in [0, 1): a fresh uniform number (the
// dispersive transmission lobe picks its colour channel with it)
static inline int findLobe(const PbrMaterial& mat, float VdotN, float rndVal, float& lobeU)
{
  float w[LOBE_COUNT];
  computeLobeWeights(mat, VdotN, w);
  int   lobe = LOBE_COUNT;
  float weight = 0.0f;
  while(--lobe > 0)
  {
    const float prev = weight;
    weight += w[lobe];
    if(rndVal < weight)
    {
      lobeU = (rndVal - prev) / w[lobe];
      break;
    }
  }
  return lobe;  // falls through to LOBE_DIFFUSE_REFLECTION
}

// KHR_materials_dispersion: one colour channel per refraction event, per-channel index of refraction as the reference's
// rasteriser defines it (shaders/gltf_raster.slang:204-208), weight 3 for the channel followed.  (nvshaders body of this
// branch is external to the reference tree: restated from the extension's definition, unpinned.)
static inline float3 applyDispersion(PbrMaterial& mat, float lobeU)
{
  const int   channel = std::min((int)(lobeU * 3.0f), 2);
  const float spread = (float)(channel - 1) * 0.025f * mat.dispersion;
  mat.ior1 = mat.ior1 + (mat.ior1 - 1.0f) * spread;
  mat.ior2 = mat.ior2 + (mat.ior2 - 1.0f) * spread;
  return f3(channel == 0 ? 3.0f : 0.0f, channel == 1 ? 3.0f : 0.0f, channel == 2 ? 3.0f : 0.0f);
}

// ---- diffuse ----------------------------------------------------------------------------------
static inline void brdf_diffuse_eval(BsdfEvaluateData& d, const PbrMaterial& mat, float3 tint)
{
  if(dot(d.k2, mat.Ng) <= 0.0f)
    return;
  d.pdf = fmaxf(0.0f, dot(d.k2, mat.N) * M_1_PI_F);
  d.bsdf_diffuse = tint * d.pdf;
}
static inline void brdf_diffuse_sample(BsdfSampleData& d, const PbrMaterial& mat, float3 tint)
{
  float3 l = cosineSampleHemisphere(d.xi.x, d.xi.y);
  d.k2 = normalize(mat.T * l.x + mat.B * l.y + mat.N * l.z);
  d.pdf = dot(d.k2, mat.N) * M_1_PI_F;
  d.bsdf_over_pdf = tint;
  d.event_type = (0.0f < dot(d.k2, mat.Ng)) ? BSDF_EVENT_DIFFUSE_REFLECTION : BSDF_EVENT_ABSORB;
}
// KHR_materials_diffuse_transmission: Lambertian lobe on the far side of the surface
static inline void btdf_diffuse_eval(BsdfEvaluateData& d, const PbrMaterial& mat, float3 tint)
{
  if(dot(d.k2, mat.Ng) >= 0.0f)
    return;
  d.pdf = fmaxf(0.0f, -dot(d.k2, mat.N) * M_1_PI_F);
  d.bsdf_diffuse = tint * d.pdf;
}
static inline void btdf_diffuse_sample(BsdfSampleData& d, const PbrMaterial& mat, float3 tint)
{
  float3 l = cosineSampleHemisphere(d.xi.x, d.xi.y);
  d.k2 = normalize(mat.T * l.x + mat.B * l.y - mat.N * l.z);
  d.pdf = -dot(d.k2, mat.N) * M_1_PI_F;
  d.bsdf_over_pdf = tint;
  d.event_type = (dot(d.k2, mat.Ng) < 0.0f) ? BSDF_EVENT_DIFFUSE_TRANSMISSION : BSDF_EVENT_ABSORB;
}

// ---- GGX-Smith reflection ---------------------------------------------------------------------
static inline void iridescenceTint(const PbrMaterial& mat, int lobe, float kh, float3& tint)
{
  if(mat.iridescence > 0.0f)
  {
    const float3 factor = thin_film_factor(mat.iridescenceThickness, mat.iridescenceIor, mat.ior2, mat.ior1, kh);
    if(lobe == LOBE_SPECULAR_REFLECTION)
      tint *= lerp3(f3(1.0f), factor, mat.iridescence);
    else if(lobe == LOBE_METAL_REFLECTION)
      tint = lerp3(tint, mat.specularColor * factor, mat.iridescence);
  }
}

// ---- bsdfEvaluateSimple / bsdfSampleSimple (nvshaders, external; call site pathtrace_functions.h.slang:537-540, the shadow catcher's
// continuation ray): Lambert + one GGX lobe with f0 = lerp(0.04, baseColor, metallic); lobe picked by the Schlick weight at N.V, value
// and pdf from evaluating both lobes.  Restated (parity unpinned, like the rest of this file). ----
static inline float schlick_f0(float f0, float vdoth)
{
  const float m = 1.0f - vdoth, m2 = m * m;
  return f0 + (1.0f - f0) * (m2 * m2 * m);
}

static inline void bsdfEvaluateSimple(BsdfEvaluateData& d, const PbrMaterial& mat)
{
  d.bsdf_diffuse = d.bsdf_glossy = f3(0.0f);
  d.pdf = 0.0f;
  const float3 h = normalize(d.k1 + d.k2);
  const float  nv = clampf(dot(mat.N, d.k1), 0.0f, 1.0f), nl = clampf(dot(mat.N, d.k2), 0.0f, 1.0f);
  const float  vh = clampf(dot(d.k1, h), 0.0f, 1.0f), nh = clampf(dot(mat.N, h), 0.0f, 1.0f);
  if(nv == 0.0f || nl == 0.0f || vh == 0.0f || nh == 0.0f)
    return;
  const float  c_min_reflectance = 0.04f;
  const float3 f0 = f3(c_min_reflectance) + (mat.baseColor - f3(c_min_reflectance)) * mat.metallic;
  const float3 fGlossy = f3(schlick_f0(f0.x, vh), schlick_f0(f0.y, vh), schlick_f0(f0.z, vh));
  const float  fDiffuse = (1.0f - mat.metallic) * (1.0f - schlick_f0(c_min_reflectance, vh));
  const float3 localH = f3(dot(mat.T, h), dot(mat.B, h), nh);
  const float  dd = hvd_ggx_eval(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), localH);
  float        G1, G2;
  ggx_smith_shadow_mask(G1, G2, f3(dot(mat.T, d.k1), dot(mat.B, d.k1), nv), f3(dot(mat.T, d.k2), dot(mat.B, d.k2), nl), mat.roughness);
  const float diffusePdf = M_1_PI_F * nl;
  const float specularPdf = G1 * dd * 0.25f / (nv * nh);
  d.pdf = specularPdf + (diffusePdf - specularPdf) * fDiffuse;
  d.bsdf_diffuse = mat.baseColor * (fDiffuse * diffusePdf);
  d.bsdf_glossy = fGlossy * (G2 * specularPdf);
}

static inline void bsdfSampleSimple(BsdfSampleData& d, const PbrMaterial& mat)
{
  d.k2 = f3(0.0f);
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  const float nv = clampf(dot(mat.N, d.k1), 0.0f, 1.0f);
  if(nv == 0.0f)
    return;
  const float fDiffuse = (1.0f - mat.metallic) * (1.0f - schlick_f0(0.04f, nv));
  int         ev;
  if(d.xi.z <= fDiffuse)
  {
    const float3 l = cosineSampleHemisphere(d.xi.x, d.xi.y);
    d.k2 = mat.T * l.x + mat.B * l.y + mat.N * l.z;
    ev = BSDF_EVENT_DIFFUSE_REFLECTION;
  }
  else
  {
    const float3 h0 = hvd_ggx_sample_vndf(f3(dot(d.k1, mat.T), dot(d.k1, mat.B), nv), mat.roughness, f2(d.xi.x, d.xi.y));
    const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
    d.k2 = h * (2.0f * dot(d.k1, h)) - d.k1;
    ev = BSDF_EVENT_GLOSSY_REFLECTION;
  }
  if(dot(d.k2, mat.N) <= 0.0f)
    return;
  BsdfEvaluateData e;
  e.k1 = d.k1;
  e.k2 = d.k2;
  e.xi = d.xi;
  bsdfEvaluateSimple(e, mat);
  const float3 total = e.bsdf_diffuse + e.bsdf_glossy;
  if(!(e.pdf > 0.00001f) || total.x != total.x || total.y != total.y || total.z != total.z)
    return;
  d.pdf = e.pdf;
  d.bsdf_over_pdf = total / e.pdf;
  d.event_type = ev;
}

static inline void brdf_ggx_smith_eval(BsdfEvaluateData& d, const PbrMaterial& mat, int lobe, float3 tint)
{
  if(dot(d.k2, mat.Ng) <= 0.0f)
    return;  // reflection lobe: nothing on the back side
  const float  nk1 = fabsf(dot(d.k1, mat.N));
  const float  nk2 = fabsf(dot(d.k2, mat.N));
  const float3 h = normalize(d.k1 + d.k2);
  const float  nh = dot(mat.N, h);
  const float  k1h = dot(d.k1, h);
  const float  k2h = dot(d.k2, h);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  const float3 h0 = f3(dot(mat.T, h), dot(mat.B, h), nh);
  d.pdf = hvd_ggx_eval(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0);
  float       G1, G2;
  const float G12 = ggx_smith_shadow_mask(G1, G2, f3(dot(mat.T, d.k1), dot(mat.B, d.k1), nk1),
                                          f3(dot(mat.T, d.k2), dot(mat.B, d.k2), nk2), mat.roughness);
  d.pdf *= 0.25f / (nk1 * nh);
  float3 bsdf = f3(G12 * d.pdf);
  d.pdf *= G1;
  iridescenceTint(mat, lobe, k1h, tint);
  d.bsdf_glossy = bsdf * tint;
}

static inline void brdf_ggx_smith_sample(BsdfSampleData& d, const PbrMaterial& mat, int lobe, float3 tint)
{
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  const float nk1 = fabsf(dot(d.k1, mat.N));
  if(nk1 <= 0.0f)
    return;
  const float3 k10 = f3(dot(d.k1, mat.T), dot(d.k1, mat.B), nk1);
  const float3 h0 = hvd_ggx_sample_vndf(k10, mat.roughness, f2(d.xi.x, d.xi.y));
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
  const float  kh = dot(d.k1, h);
  if(kh <= 0.0f)
    return;
  d.k2 = h * (2.0f * kh) - d.k1;
  if(dot(d.k2, mat.Ng) <= 0.0f)
    return;
  const float nk2 = fabsf(dot(d.k2, mat.N));
  float       G1, G2;
  const float G12 = ggx_smith_shadow_mask(G1, G2, k10, f3(dot(d.k2, mat.T), dot(d.k2, mat.B), nk2), mat.roughness);
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = f3(G12 / G1);
  d.pdf = hvd_ggx_eval(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0) * G1;
  d.pdf *= 0.25f / (nk1 * h0.z);
  iridescenceTint(mat, lobe, kh, tint);
  d.bsdf_over_pdf *= tint;
  d.event_type = BSDF_EVENT_GLOSSY_REFLECTION;
}

// ---- GGX-Smith transmission -------------------------------------------------------------------
static inline void btdf_ggx_smith_eval(BsdfEvaluateData& d, const PbrMaterial& mat, float3 tint)
{
  const bool  thin = (mat.thickness == 0.0f);
  const float nk1 = fabsf(dot(d.k1, mat.N));
  const float nk2 = fabsf(dot(d.k2, mat.N));
  const bool  backside = (dot(d.k2, mat.Ng) <= 0.0f);
  const float3 h = compute_half_vector(d.k1, d.k2, mat.N, mat.ior1, mat.ior2, nk2, backside, thin);
  const float  nh = dot(mat.N, h);
  const float  k1h = dot(d.k1, h);
  // thin-walled pseudo-BTDF: the half vector pairs k1 with the MIRRORED k2, so that is the direction whose
  // cosine to h must be checked (keeps bsdfEvaluate consistent with what bsdfSample generates)
  const float  k2h = (backside && thin) ? dot(d.k2 + mat.N * (nk2 + nk2), h) : dot(d.k2, h) * (backside ? -1.0f : 1.0f);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  float fr;
  if(!backside)
  {
    if(!isTIR(mat.ior1, mat.ior2, k1h))
      return;  // only total internal reflection reflects in the pure-transmission lobe
    fr = 1.0f;
  }
  else
    fr = 0.0f;
  const float3 h0 = f3(dot(mat.T, h), dot(mat.B, h), nh);
  d.pdf = hvd_ggx_eval(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0);
  float       G1, G2;
  const float G12 = ggx_smith_shadow_mask(G1, G2, f3(dot(mat.T, d.k1), dot(mat.B, d.k1), nk1),
                                          f3(dot(mat.T, d.k2), dot(mat.B, d.k2), nk2), mat.roughness);
  if(!thin && backside)
  {
    const float tmp = k1h * mat.ior1 - k2h * mat.ior2;
    d.pdf *= k1h * k2h / (nk1 * nh * tmp * tmp);
  }
  else
    d.pdf *= 0.25f / (nk1 * nh);
  const float  prob = backside ? 1.0f - fr : fr;
  const float3 bsdf = f3(prob * G12 * d.pdf);
  d.pdf *= prob * G1;
  d.bsdf_glossy = bsdf * tint;
}

static inline void btdf_ggx_smith_sample(BsdfSampleData& d, const PbrMaterial& mat, float3 tint)
{
  const bool thin = (mat.thickness == 0.0f);
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  const float  nk1 = fabsf(dot(d.k1, mat.N));
  const float3 k10 = f3(dot(d.k1, mat.T), dot(d.k1, mat.B), nk1);
  const float3 h0 = hvd_ggx_sample_vndf(k10, mat.roughness, f2(d.xi.x, d.xi.y));
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
  const float  kh = dot(d.k1, h);
  if(kh <= 0.0f)
    return;
  bool tir = false;
  if(thin)
  {
    float3 r = h * (2.0f * kh) - d.k1;
    d.k2 = normalize(r - mat.N * (2.0f * dot(r, mat.N)));
  }
  else
    d.k2 = refract_h(d.k1, h, mat.ior1 / mat.ior2, kh, tir);
  const int   ev = tir ? BSDF_EVENT_GLOSSY_REFLECTION : BSDF_EVENT_GLOSSY_TRANSMISSION;
  const float gnk2 = dot(d.k2, mat.Ng) * ((ev == BSDF_EVENT_GLOSSY_REFLECTION) ? 1.0f : -1.0f);
  if(gnk2 <= 0.0f)
    return;
  const float nk2 = fabsf(dot(d.k2, mat.N));
  const float k2h = fabsf(dot(d.k2, h));
  float       G1, G2;
  const float G12 = ggx_smith_shadow_mask(G1, G2, k10, f3(dot(d.k2, mat.T), dot(d.k2, mat.B), nk2), mat.roughness);
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = f3(G12 / G1);
  d.pdf = hvd_ggx_eval(f2(1.0f / mat.roughness.x, 1.0f / mat.roughness.y), h0) * G1;
  if(!thin && ev == BSDF_EVENT_GLOSSY_TRANSMISSION)
  {
    const float tmp = kh * mat.ior1 - k2h * mat.ior2;
    d.pdf *= kh * k2h / (nk1 * h0.z * tmp * tmp);
  }
  else
    d.pdf *= 0.25f / (nk1 * h0.z);
  d.bsdf_over_pdf *= tint;
  d.event_type = ev;
}

// ---- sheen ------------------------------------------------------------------------------------
static inline void brdf_sheen_eval(BsdfEvaluateData& d, const PbrMaterial& mat)
{
  if(dot(d.k2, mat.Ng) <= 0.0f)
    return;
  const float  nk1 = fabsf(dot(d.k1, mat.N));
  const float  nk2 = fabsf(dot(d.k2, mat.N));
  const float3 h = normalize(d.k1 + d.k2);
  const float  nh = dot(mat.N, h);
  const float  k1h = dot(d.k1, h);
  const float  k2h = dot(d.k2, h);
  if(nk1 <= 0.0f || nh <= 0.0f || k1h < 0.0f || k2h < 0.0f)
    return;
  const float invRoughness = 1.0f / (mat.sheenRoughness * mat.sheenRoughness);
  d.pdf = hvd_sheen_eval(invRoughness, nh);
  float       G1, G2;
  const float G12 = vcavities_shadow_mask(G1, G2, nh, f3(dot(mat.T, d.k1), dot(mat.B, d.k1), nk1), k1h,
                                          f3(dot(mat.T, d.k2), dot(mat.B, d.k2), nk2), k2h);
  d.pdf *= 0.25f / (nk1 * nh);
  const float3 bsdf = f3(d.pdf * G12);
  d.pdf *= G1;
  d.bsdf_glossy = bsdf * mat.sheenColor;
}
static inline void brdf_sheen_sample(BsdfSampleData& d, const PbrMaterial& mat)
{
  d.bsdf_over_pdf = f3(0.0f);
  d.pdf = 0.0f;
  d.event_type = BSDF_EVENT_ABSORB;
  const float nk1 = fabsf(dot(d.k1, mat.N));
  if(nk1 <= 0.0f)
    return;
  const float3 k10 = f3(dot(d.k1, mat.T), dot(d.k1, mat.B), nk1);
  const float  invRoughness = 1.0f / (mat.sheenRoughness * mat.sheenRoughness);
  // xi.z was consumed by the lobe choice; re-use xi.x's low bits is not possible in fp32, so the
  // flip decision takes xi.z (uniform, independent of xi.xy)
  const float3 h0 = flip(hvd_sheen_sample(f2(d.xi.x, d.xi.y), invRoughness), k10, d.xi.z);
  if(fabsf(h0.z) == 0.0f)
    return;
  const float3 h = mat.T * h0.x + mat.B * h0.y + mat.N * h0.z;
  const float  k1h = dot(d.k1, h);
  if(k1h <= 0.0f)
    return;
  d.k2 = h * (2.0f * k1h) - d.k1;
  if(dot(d.k2, mat.Ng) <= 0.0f)
    return;
  const float nk2 = fabsf(dot(d.k2, mat.N));
  const float k2h = fabsf(dot(d.k2, h));
  float       G1, G2;
  const float G12 = vcavities_shadow_mask(G1, G2, h0.z, k10, k1h, f3(dot(d.k2, mat.T), dot(d.k2, mat.B), nk2), k2h);
  if(G12 <= 0.0f)
    return;
  d.bsdf_over_pdf = mat.sheenColor * (G12 / G1);
  d.pdf = hvd_sheen_eval(invRoughness, h0.z) * G1;
  d.pdf *= 0.25f / (nk1 * h0.z);
  d.event_type = BSDF_EVENT_GLOSSY_REFLECTION;
}

// clearcoat lobe: isotropic GGX about the clearcoat normal, no iridescence
static inline void setupClearcoat(PbrMaterial& mat)
{
  const float a = mat.clearcoatRoughness * mat.clearcoatRoughness;
  mat.roughness = f2(a, a);
  mat.N = mat.Nc;
  mat.B = normalize(cross(mat.N, mat.T));
  mat.T = cross(mat.B, mat.N);
  mat.iridescence = 0.0f;
}

// ---- top level (call sites: gltf_pathtrace.slang:337,362) --------------------------------------
static inline void bsdfEvaluate(BsdfEvaluateData& d, const PbrMaterial& matIn)
{
  PbrMaterial mat = matIn;
  const float VdotN = dot(d.k1, mat.N);
  float       lobeU = 0.0f;
  const int   lobe = findLobe(mat, VdotN, d.xi.z, lobeU);
  d.bsdf_diffuse = f3(0.0f);
  d.bsdf_glossy = f3(0.0f);
  d.pdf = 0.0f;
  switch(lobe)
  {
    case LOBE_DIFFUSE_REFLECTION:
      brdf_diffuse_eval(d, mat, mat.baseColor);
      break;
    case LOBE_DIFFUSE_TRANSMISSION:
      btdf_diffuse_eval(d, mat, mat.diffuseTransmissionColor);
      break;
    case LOBE_SPECULAR_REFLECTION:
      brdf_ggx_smith_eval(d, mat, LOBE_SPECULAR_REFLECTION, mat.specularColor);
      break;
    case LOBE_SPECULAR_TRANSMISSION:
      if(mat.dispersion > 0.0f)
      {
        const float3 base = mat.baseColor;
        const float3 w = applyDispersion(mat, lobeU);
        btdf_ggx_smith_eval(d, mat, base * w);
      }
      else
        btdf_ggx_smith_eval(d, mat, mat.baseColor);
      break;
    case LOBE_METAL_REFLECTION:
      brdf_ggx_smith_eval(d, mat, LOBE_METAL_REFLECTION, mat.baseColor);
      break;
    case LOBE_CLEARCOAT_REFLECTION:
      setupClearcoat(mat);
      brdf_ggx_smith_eval(d, mat, LOBE_CLEARCOAT_REFLECTION, f3(1.0f));
      break;
    case LOBE_SHEEN_REFLECTION:
      brdf_sheen_eval(d, mat);
      break;
  }
}

static inline void bsdfSample(BsdfSampleData& d, const PbrMaterial& matIn)
{
  PbrMaterial mat = matIn;
  const float VdotN = dot(d.k1, mat.N);
  float       lobeU = 0.0f;
  const int   lobe = findLobe(mat, VdotN, d.xi.z, lobeU);
  d.pdf = 0.0f;
  d.bsdf_over_pdf = f3(0.0f);
  d.event_type = BSDF_EVENT_ABSORB;
  d.k2 = f3(0.0f);
  switch(lobe)
  {
    case LOBE_DIFFUSE_REFLECTION:
      brdf_diffuse_sample(d, mat, mat.baseColor);
      break;
    case LOBE_DIFFUSE_TRANSMISSION:
      btdf_diffuse_sample(d, mat, mat.diffuseTransmissionColor);
      break;
    case LOBE_SPECULAR_REFLECTION:
      brdf_ggx_smith_sample(d, mat, LOBE_SPECULAR_REFLECTION, mat.specularColor);
      break;
    case LOBE_SPECULAR_TRANSMISSION:
      if(mat.dispersion > 0.0f)
      {
        const float3 base = mat.baseColor;
        const float3 w = applyDispersion(mat, lobeU);
        btdf_ggx_smith_sample(d, mat, base * w);
      }
      else
        btdf_ggx_smith_sample(d, mat, mat.baseColor);
      break;
    case LOBE_METAL_REFLECTION:
      brdf_ggx_smith_sample(d, mat, LOBE_METAL_REFLECTION, mat.baseColor);
      break;
    case LOBE_CLEARCOAT_REFLECTION:
      setupClearcoat(mat);
      brdf_ggx_smith_sample(d, mat, LOBE_CLEARCOAT_REFLECTION, f3(1.0f));
      break;
    case LOBE_SHEEN_REFLECTION:
      brdf_sheen_sample(d, mat);
      break;
  }
  if(d.pdf <= 0.00001f || std::isnan(d.bsdf_over_pdf.x) || std::isnan(d.bsdf_over_pdf.y) || std::isnan(d.bsdf_over_pdf.z))
    d.event_type = BSDF_EVENT_ABSORB;
}

// ---------------------------------------------------------------------------------------------
// misc nvshaders functions
// ---------------------------------------------------------------------------------------------
static inline float3 pointOffset(float3 p, float3 p_a, float3 p_b, float3 p_c, float3 n_a, float3 n_b, float3 n_c, float3 bary)
{
  float3      tmpu = p - p_a, tmpv = p - p_b, tmpw = p - p_c;
  const float dotu = fminf(0.0f, dot(tmpu, n_a));
  const float dotv = fminf(0.0f, dot(tmpv, n_b));
  const float dotw = fminf(0.0f, dot(tmpw, n_c));
  tmpu -= n_a * dotu;
  tmpv -= n_b * dotv;
  tmpw -= n_c * dotw;
  return p + (tmpu * bary.x + tmpv * bary.y + tmpw * bary.z);
}

static inline float4 makeFastTangent(float3 n)
{
  if(n.z < -0.99998796f)
    return f4(0.0f, -1.0f, 0.0f, 1.0f);
  const float a = 1.0f / (1.0f + n.z);
  const float b = -n.x * n.y * a;
  return f4(1.0f - n.x * n.x * a, b, -n.x, 1.0f);
}

static inline float2 getSphericalUv(float3 v)
{
  const float gamma = asinf(-v.y);
  const float theta = atan2f(v.z, v.x);
  return f2(theta * (M_1_PI_F * 0.5f) + 0.5f, gamma * M_1_PI_F + 0.5f);
}

static inline float3 rotate(float3 v, float3 k, float theta)
{
  const float c = cosf(theta), s = sinf(theta);
  return (v * c) + (cross(k, v) * s) + (k * dot(k, v)) * (1.0f - c);
}

static inline float henyeyGreensteinPdf(float cosTheta, float g)
{
  const float denom = 1.0f + g * g - 2.0f * g * cosTheta;
  return (1.0f / (4.0f * M_PI_F)) * (1.0f - g * g) / (denom * sqrtf(denom));
}
// samples the new propagation direction around wi (forward = +g)
static inline float3 sampleHenyeyGreenstein(float2 xi, float g, float3 wi)
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
  const float  phi = M_TWO_PI_F * xi.y;
  const float4 t = makeFastTangent(wi);
  const float3 T = normalize(xyz(t));
  const float3 B = cross(wi, T);
  return normalize(T * (sinTheta * cosf(phi)) + B * (sinTheta * sinf(phi)) + wi * cosTheta);
}

}  // namespace orc
