// headless_main.cpp — stand-in for `vk_gltf_renderer --headless --frames N --ptSamples S …` (reference src/main.cpp:133-136,
// headless flow src/renderer.cpp:1939-1977, summary lines src/benchmarking.cpp:245-305) on top of the C++ host mirror.
//
//   b200pt_headless --scene scene.b2sc [--size W H] [--frames N] [--warmupFrames K] [--ptMaxDepth D] [--ptSamples S] …
//                   [--device G] [--framesInFlight L] [--out image.pfm]
//
// Prints the reference's two record kinds so its benchmark tooling can read them: a human-readable
// "HEADLESS_SUMMARY key=value …" line and a "BENCHMARK_JSON {…}" line with the same keys, plus the ray counters the
// reference lacks.  The image written by --out is the RGBA32F accumulation buffer (what the reference tonemaps) as PFM.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "b200pt_host.hpp"

using namespace b200host;

static void writePfm(const std::string& path, const std::vector<float>& rgba, int w, int h)
{
  FILE* f = std::fopen(path.c_str(), "wb");
  if(!f)
    throw Error("cannot write " + path);
  std::fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
  std::vector<float> row((size_t)w * 3);
  for(int y = h - 1; y >= 0; y--)  // PFM stores the bottom row first
  {
    for(int x = 0; x < w; x++)
      for(int c = 0; c < 3; c++)
        row[(size_t)x * 3 + c] = rgba[((size_t)y * w + x) * 4 + c];
    std::fwrite(row.data(), sizeof(float), row.size(), f);
  }
  std::fclose(f);
}

static void writeRaw(const std::string& path, const std::vector<float>& rgba)
{
  FILE* f = std::fopen(path.c_str(), "wb");
  if(!f)
    throw Error("cannot write " + path);
  std::fwrite(rgba.data(), sizeof(float), rgba.size(), f);
  std::fclose(f);
}

int main(int argc, char** argv)
{
  std::string scenePath, outPath, rawPath;
  int         width = 1920, height = 1080, frames = 16, warmupFrames = 0, framesInFlight = 0;
  Resources   res;
  PathTracer  pt;
  try
  {
    for(int i = 1; i < argc; i++)
    {
      const std::string a = argv[i];
      auto              next = [&]() -> std::string {
        if(i + 1 >= argc)
          throw Error("missing value after " + a);
        return argv[++i];
      };
      if(a == "--scene")
        scenePath = next();
      else if(a == "--size")
      {
        width = std::stoi(next());
        height = std::stoi(next());
      }
      else if(a == "--frames")
        frames = std::stoi(next());
      else if(a == "--warmupFrames")
        warmupFrames = std::stoi(next());
      else if(a == "--device")
        res.cudaDevice = std::stoi(next());
      else if(a == "--framesInFlight")
        framesInFlight = std::stoi(next());
      else if(a == "--maxFrames")
        res.settings.maxFrames = std::stoi(next());
      else if(a == "--hdrEnvIntensity")
        res.settings.hdrEnvIntensity = std::stof(next());
      else if(a == "--hdrEnvRotation")
        res.settings.hdrEnvRotation = std::stof(next());
      else if(a == "--envSystem")
        res.settings.envSystem = std::stoi(next());
      else if(a == "--out")
        outPath = next();
      else if(a == "--outRaw")
        rawPath = next();
      else if(a.rfind("--pt", 0) == 0)
      {
        if(!pt.setParameter(a.substr(2), next()))
          throw Error("unknown parameter " + a);
      }
      else
        throw Error("unknown argument " + a);
    }
    if(scenePath.empty())
      throw Error("usage: b200pt_headless --scene scene.b2sc [--size W H] [--frames N] [--pt<Name> value] [--out image.pfm]");

    SceneData scene;
    scene.load(scenePath);
    res.scene = &scene;
    res.camera = scene.camera;
    res.width = width;
    res.height = height;
    if(!scene.hdrRgb.empty())
    {
      res.hdrRgb = scene.hdrRgb.data();
      res.hdrWidth = scene.hdrWidth;
      res.hdrHeight = scene.hdrHeight;
    }
    pt.onAttach(res);
    if(framesInFlight > 0)
      pt.setFramesInFlight(framesInFlight);

    // frame loop with the reference's warm-up split (benchmarking.cpp: measured timer starts after warmupFrames)
    using clock = std::chrono::steady_clock;
    const auto t0 = clock::now();
    auto       tMeasured = t0;
    res.frameCount = -1;
    for(int f = 0; f < frames; f++)
    {
      if(f == warmupFrames)
      {
        pt.synchronize();
        pt.resetStats();
        tMeasured = clock::now();
      }
      res.frameCount++;
      pt.onRender(res);
    }
    pt.synchronize();
    const auto   t1 = clock::now();
    const double totalWallMs = std::chrono::duration<double, std::milli>(t1 - t0).count();
    const double measuredWallMs = std::chrono::duration<double, std::milli>(t1 - tMeasured).count();
    const int    measuredFrames = frames - warmupFrames;
    const int    accumFrames = std::min(frames, std::max(res.settings.maxFrames, 0));
    const int    effectiveSpp = accumFrames * std::max(pt.ptSamples, 1);
    const int    measuredEffectiveSpp = std::max(0, std::min(accumFrames - warmupFrames, measuredFrames)) * std::max(pt.ptSamples, 1);
    const double sec = measuredWallMs / 1000.0;
    const double throughputMSps = sec > 0 ? (double)width * height * measuredEffectiveSpp / sec / 1e6 : 0.0;
    const double sppPerSec = sec > 0 ? measuredEffectiveSpp / sec : 0.0;
    const b200pt_stats st = pt.stats();
    const double       mrays = sec > 0 ? (double)(st.closestRays + st.shadowRays) / sec / 1e6 : 0.0;

    std::printf(
        "HEADLESS_SUMMARY frames=%d maxFrames=%d ptSamples=%d effective_spp=%d measured_effective_spp=%d resolution=%dx%d wall_ms=%.3f "
        "ms_per_frame=%.3f total_wall_ms=%.3f total_ms_per_frame=%.3f warmup_frames=%d measured_frames=%d throughput_MSps=%.3f spp_per_sec=%.2f\n",
        frames, res.settings.maxFrames, pt.ptSamples, effectiveSpp, measuredEffectiveSpp, width, height, measuredWallMs,
        measuredFrames > 0 ? measuredWallMs / measuredFrames : 0.0, totalWallMs, frames > 0 ? totalWallMs / frames : 0.0, warmupFrames, measuredFrames,
        throughputMSps, sppPerSec);
    std::printf(
        "BENCHMARK_JSON {\"schema\":1,\"type\":\"headless_summary\",\"frames\":%d,\"maxFrames\":%d,\"ptSamples\":%d,\"effective_spp\":%d,"
        "\"measured_effective_spp\":%d,\"resolution_w\":%d,\"resolution_h\":%d,\"wall_ms\":%.3f,\"ms_per_frame\":%.3f,\"total_wall_ms\":%.3f,"
        "\"total_ms_per_frame\":%.3f,\"warmup_frames\":%d,\"measured_frames\":%d,\"throughput_MSps\":%.3f,\"spp_per_sec\":%.2f,"
        "\"closest_rays\":%llu,\"shadow_rays\":%llu,\"Mray_per_s\":%.3f,\"triangles\":%zu,\"backend\":\"b200pt\"}\n",
        frames, res.settings.maxFrames, pt.ptSamples, effectiveSpp, measuredEffectiveSpp, width, height, measuredWallMs,
        measuredFrames > 0 ? measuredWallMs / measuredFrames : 0.0, totalWallMs, frames > 0 ? totalWallMs / frames : 0.0, warmupFrames, measuredFrames,
        throughputMSps, sppPerSec, (unsigned long long)st.closestRays, (unsigned long long)st.shadowRays, mrays, scene.triangleCount());

    if(!outPath.empty() || !rawPath.empty())
    {
      const std::vector<float> img = pt.readAccum();
      if(!outPath.empty())
        writePfm(outPath, img, width, pt.tileRows());
      if(!rawPath.empty())
        writeRaw(rawPath, img);
    }
    pt.onDetach(res);
  }
  catch(const std::exception& e)
  {
    std::fprintf(stderr, "b200pt_headless: %s\n", e.what());
    return 1;
  }
  return 0;
}
