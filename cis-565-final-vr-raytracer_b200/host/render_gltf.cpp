// render_gltf.cpp — headless C++ harness: what SampleExample::loadScene + the main loop do around the hot path
// (sample_example.cpp:82-92, 164-196, 376-415), minus window/GUI.  Usage:
//   render_gltf scene.gltf out.pfm [width height frames [environment.hdr]]
// Build: g++ -std=c++17 -I include -I cis-565-final-vr-raytracer_b200/host render_gltf.cpp -L<dir> -leidola
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "eidola.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s scene.gltf out.pfm [w h frames [environment.hdr]]\n", argv[0]); return 2; }
  const uint32_t w = argc > 3 ? std::atoi(argv[3]) : 1920, h = argc > 4 ? std::atoi(argv[4]) : 1080;
  const int frames = argc > 5 ? std::atoi(argv[5]) : 16;
  try {
    eidola::Scene scene;
    scene.setup(0);
    if (!scene.load(argv[1])) { std::fprintf(stderr, "load failed: %s\n", eid_last_error()); return 1; }
    eidola::AccelStructure accel;
    accel.create(scene);
    eidola::Renderer renderer;
    renderer.create({w, h}, scene, accel);
    const float env[3] = {0.25f, 0.25f, 0.25f};
    renderer.setEnvironmentConstant(env);
    const eid_scene_info info = scene.getStat();
    RtxState st = eidola::defaultRtxState(w, h);
    eidola::HdrSampling skydome;
    if (argc > 6) {                                             // SampleExample::loadEnvironmentHdr (sample_example.cpp:97-106)
      skydome.setup(0);
      skydome.loadEnvironment(argv[6]);
      renderer.setEnvironment(&skydome);
      st.fireflyClampThreshold = skydome.getIntegral() * 4.f;   // "magic" there too; environmentProb stays at its default 0.25
      st.envMapLuminIntegInv = 1.f / skydome.getIntegral();
    } else {
      st.environmentProb = 0.f;                                 // constant environment: nothing to importance-sample
      st.fireflyClampThreshold = 4.f * 3.14159265f;             // integral of the constant environment * 4 (sample_example.cpp:104)
      st.envMapLuminIntegInv = 1.f / 3.14159265f;
    }
    st.lightLuminIntegInv = 1.f / (info.trigLightWeight + info.puncLightWeight);   // sample_example.cpp:87
    std::vector<float> direct((size_t)w * h * 4), indirect((size_t)w * h * 4);
    scene.updateCamera({w, h});
    for (int f = 0; f < frames; ++f) {
      scene.updateCamera({w, h});
      st.frame = f; st.time = 1000u + 16u * f;
      if (f + 1 < frames) renderer.run(st, f);
      else renderer.renderToHost(nullptr, st, f, direct.data(), indirect.data());
    }
    const eid_frame_stats s = renderer.stats();
    std::printf("rays/frame: %llu closest + %llu any\n", (unsigned long long)s.closestHitRays, (unsigned long long)s.anyHitRays);
    FILE* fp = std::fopen(argv[2], "wb");                       // PFM, bottom-up rows, direct+indirect summed like post.frag:129
    if (!fp) return 1;
    std::fprintf(fp, "PF\n%u %u\n-1.0\n", w, h);
    for (int y = (int)h - 1; y >= 0; --y)
      for (uint32_t x = 0; x < w; ++x) {
        size_t i = ((size_t)y * w + x) * 4;
        float rgb[3] = {direct[i] + indirect[i], direct[i + 1] + indirect[i + 1], direct[i + 2] + indirect[i + 2]};
        std::fwrite(rgb, 4, 3, fp);
      }
    std::fclose(fp);
  } catch (const eidola::Error& e) { std::fprintf(stderr, "eidola error %d: %s\n", e.code, e.what()); return 1; }
  return 0;
}
