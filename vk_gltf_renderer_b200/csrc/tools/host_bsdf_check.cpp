// host_bsdf_check.cpp — the device BSDF source (../bsdf.cuh: bsdfEvaluate / bsdfSample, every lobe, FEAT_ALL and FEAT_LEAN)
// compiled for the host through host_shim.h and compared with the oracle's outputs for the same records.
//
//   host_bsdf_check records.bin        records.bin = u32 n, n x 48 floats (bsdf_io.py packing), n x 8 oracle eval, n x 8 oracle sample
//                                      [, n x 8 oracle bsdfSampleSimple: the shadow catcher's continuation BSDF]
//
// Both sides then use the same libm, so what is compared is the SOURCE: oracle/bsdf.h vs csrc/bsdf.cuh, operation by
// operation (build with -ffp-contract=off, the host analogue of the kernels' -fmad=false).  On the GPU the only remaining
// difference is CUDA's libm (tests/test_gpu_scenes.py::test_bsdf_parity allows 2e-4 for it).
#include "host_shim.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../bsdf.cuh"

using namespace pt;

// same packing as unpackTestMaterial in ../b200pt.cu / vk_gltf_renderer_b200/bsdf_io.py
static PbrMaterial unpack(const float* p)
{
  PbrMaterial m;
  m.baseColor = f3(p[0], p[1], p[2]);
  m.opacity = 1.0f;
  m.roughness = f2(p[3], p[4]);
  m.metallic = p[5];
  m.emissive = f3(0.0f);
  m.N = f3(p[6], p[7], p[8]);
  m.T = f3(p[9], p[10], p[11]);
  m.B = f3(p[12], p[13], p[14]);
  m.Ng = f3(p[15], p[16], p[17]);
  m.ior1 = p[18];
  m.ior2 = p[19];
  m.specular = p[20];
  m.specularColor = f3(p[21], p[22], p[23]);
  m.transmission = p[24];
  m.attenuationColor = f3(1.0f);
  m.attenuationDistance = 1.0f;
  m.thickness = p[25];
  m.clearcoat = p[26];
  m.clearcoatRoughness = p[27];
  m.Nc = m.N;
  m.iridescence = p[28];
  m.iridescenceIor = p[29];
  m.iridescenceThickness = p[30];
  m.sheenColor = f3(p[31], p[32], p[33]);
  m.sheenRoughness = p[34];
  m.diffuseTransmissionFactor = p[35];
  m.diffuseTransmissionColor = f3(p[36], p[37], p[38]);
  m.scatterCoefficient = f3(0.0f);
  m.scatterAnisotropy = 0.0f;
  m.dispersion = 0.0f;
  m.retroreflection = 0.0f;
  return m;
}

static bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

int main(int argc, char** argv)
{
  if(argc < 2) { std::fprintf(stderr, "usage: host_bsdf_check records.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  uint32_t n = 0;
  if(!f || std::fread(&n, 4, 1, f) != 1) return 2;
  std::vector<float> rec((size_t)n * 48), oe((size_t)n * 8), os((size_t)n * 8);
  if(std::fread(rec.data(), 4, rec.size(), f) != rec.size() || std::fread(oe.data(), 4, oe.size(), f) != oe.size() || std::fread(os.data(), 4, os.size(), f) != os.size()) return 2;
  std::vector<float> oq((size_t)n * 8);
  const bool         haveSimple = std::fread(oq.data(), 4, oq.size(), f) == oq.size();
  std::fclose(f);
  uint64_t badE = 0, badS = 0, badQ = 0, events[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double   worst = 0;
  for(uint32_t i = 0; i < n; i++)
  {
    const float*      p = &rec[(size_t)i * 48];
    const PbrMaterial m = unpack(p);
    const BsdfEval    e = bsdfEvaluate<FEAT_ALL>(m, f3(p[39], p[40], p[41]), f3(p[42], p[43], p[44]), f3(p[45], p[46], p[47]));
    const float       ge[7] = {e.bsdf_diffuse.x, e.bsdf_diffuse.y, e.bsdf_diffuse.z, e.bsdf_glossy.x, e.bsdf_glossy.y, e.bsdf_glossy.z, e.pdf};
    bool              ok = true;
    for(int k = 0; k < 7; k++)
    {
      ok = ok && same(ge[k], oe[(size_t)i * 8 + k]);
      worst = std::fmax(worst, std::fabs((double)ge[k] - oe[(size_t)i * 8 + k]) / std::fmax(1e-6, std::fabs((double)oe[(size_t)i * 8 + k])));
    }
    badE += !ok;
    const BsdfSample s = bsdfSample<FEAT_ALL>(m, f3(p[39], p[40], p[41]), f3(p[45], p[46], p[47]));
    const float      gs[8] = {s.k2.x, s.k2.y, s.k2.z, s.bsdf_over_pdf.x, s.bsdf_over_pdf.y, s.bsdf_over_pdf.z, s.pdf, (float)s.event_type};
    ok = true;
    for(int k = 0; k < 8; k++)
      ok = ok && same(gs[k], os[(size_t)i * 8 + k]);
    badS += !ok;
    events[(int)gs[7] & 7]++;
    if(haveSimple)
    {
      const BsdfSample q = bsdfSampleSimple(m, f3(p[39], p[40], p[41]), f3(p[45], p[46], p[47]));
      const float      gq[8] = {q.k2.x, q.k2.y, q.k2.z, q.bsdf_over_pdf.x, q.bsdf_over_pdf.y, q.bsdf_over_pdf.z, q.pdf, (float)q.event_type};
      ok = true;
      for(int k = 0; k < 8; k++)
        ok = ok && same(gq[k], oq[(size_t)i * 8 + k]);
      badQ += !ok;
    }
  }
  if(haveSimple)
    std::printf("bsdfSampleSimple mismatches %llu (bit level)\n", (unsigned long long)badQ);
  std::printf("records %u | eval mismatches %llu, sample mismatches %llu (bit level), worst eval relative difference %.3g\n", n, (unsigned long long)badE,
              (unsigned long long)badS, worst);
  return (badE || badS || badQ) ? 1 : 0;
}
