// vt_group_demo.cpp -- a C++ program (no Python) that renders BASELINE config 2 on N GPUs with the host classes of
// voxeltoy_b200/host/vt_host.h: RendererGroup = one Renderer per CUDA device, combined through the C ABI's vt_group
// (NCCL over NVLink; the peer-memory kernel when replicas share a device).
//
//   vt_group_demo --vox scene_fall.vox [--env env.pfm] --devices 0,1,2,3,4,5,6,7 --mode samples|tiles
//                 [--width 1920 --height 1080 --bounces 4 --passes 256 --steps 3] [--out frame.pfm] [--exchange nccl|peer]
//
// Prints one JSON line: Msamples/s over the timed steps (wall clock around renderPasses + combine, after one warm-up step).
// Built by `python -m voxeltoy_b200.build` into voxeltoy_b200/vt_group_demo (links libvoxeltoy_b200.so).
#include "../voxeltoy_b200/host/vt_host.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
    std::string vox, env, out, mode = "samples", devices = "0", exchange;
    int W = 1920, H = 1080, bounces = 4, passes = 256, steps = 3;
    float theta = 120.0f, phi = 30.0f, fstop = 2.8f;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i], v = argv[i + 1];
        if (k == "--vox") vox = v; else if (k == "--env") env = v; else if (k == "--out") out = v; else if (k == "--mode") mode = v;
        else if (k == "--devices") devices = v; else if (k == "--exchange") exchange = v;
        else if (k == "--width") W = atoi(v.c_str()); else if (k == "--height") H = atoi(v.c_str());
        else if (k == "--bounces") bounces = atoi(v.c_str()); else if (k == "--passes") passes = atoi(v.c_str());
        else if (k == "--steps") steps = atoi(v.c_str());
        else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
    }
    if (vox.empty()) { fprintf(stderr, "usage: vt_group_demo --vox file.vox [--env file.pfm] --devices 0,1 --mode samples|tiles ...\n"); return 2; }
    std::vector<int> devs;
    for (size_t p = 0; p <= devices.size();) { const size_t q = devices.find(',', p); devs.push_back(atoi(devices.substr(p, q - p).c_str())); if (q == std::string::npos) break; p = q + 1; }

    RendererGroup g;
    if (!g.initialize(devs, mode == "tiles" ? RendererGroup::MODE_TILES : RendererGroup::MODE_SAMPLES)) { fprintf(stderr, "%s\n", g.getStatus().c_str()); return 1; }
    if (!exchange.empty() && vt_group_set_exchange(g.group(), exchange == "peer" ? VT_EXCHANGE_PEER : VT_EXCHANGE_NCCL) != VT_OK) {
        fprintf(stderr, "%s\n", vt_group_last_error(g.group())); return 1;
    }
    // the set-up bench.py makes through the same classes: frame, scene, environment, thin lens, orbit, autofocus on the centre
    g.resizeFrame(W, H);
    g.loadVoxFile(vox);
    g.forEach([&](Renderer& r) {
        r.renderSettings().m_pathtracerMaxNumBounces = bounces;
        r.renderSettings().m_backgroundImage = env;
        r.updateRenderSettings();
        r.camera().setLensModel(CameraParameters::CLM_THIN_LENS);
        r.camera().controller().orbitAroundTarget(theta * 3.14159265358979f / 180.0f, phi * 3.14159265358979f / 180.0f);
        r.camera().setFStop(fstop);
        r.resetRender();
        const int32_t none[4] = { -1, -1, -1, 0 }; const float nrm[4] = { 1, 0, 0, 0 };
        vt_set_selection(r.context(), none, nrm);
    });
    g.requestAction(0.5f, 0.5f, 0.0f, 0.0f, Action::PA_SELECT_FOCAL_POINT, true);
    std::vector<float> frame((size_t)W * H * 4);
    g.renderPasses(passes);                        // warm-up step: sizes the pools, runs the autofocus action
    if (!g.readAverage(&frame[0])) { fprintf(stderr, "%s\n", g.getStatus().c_str()); return 1; }
    g.resetRender();
    vt_group_sync(g.group());
    const auto t0 = std::chrono::steady_clock::now();
    float exchange_ms = 0.f;
    for (int s = 0; s < steps; ++s) {
        g.renderPasses(passes);
        if (!g.readAverage(&frame[0])) { fprintf(stderr, "%s\n", g.getStatus().c_str()); return 1; }
        float ms = 0.f; vt_group_last_exchange_ms(g.group(), &ms); exchange_ms += ms;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double samples = (double)W * H * passes * steps * (mode == "tiles" ? 1.0 : (double)devs.size());
    if (!out.empty()) {
        std::vector<float> rgb((size_t)W * H * 3);
        for (size_t i = 0; i < (size_t)W * H; ++i) for (int c = 0; c < 3; ++c) rgb[3 * i + c] = frame[4 * i + c];
        if (!writePFM(out, &rgb[0], (unsigned)W, (unsigned)H)) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    }
    printf("{\"program\": \"vt_group_demo\", \"gpus\": %d, \"mode\": \"%s\", \"exchange\": \"%s\", \"nccl\": %d, \"width\": %d, \"height\": %d, "
           "\"bounces\": %d, \"passes_per_step\": %d, \"steps\": %d, \"seconds\": %.6f, \"msamples_per_s\": %.2f, "
           "\"exchange_ms_per_step\": %.4f, \"exchange_bytes\": %zu}\n",
           (int)devs.size(), mode.c_str(), vt_group_get_exchange(g.group()) == VT_EXCHANGE_PEER ? "peer" : "nccl", vt_nccl_version(), W, H,
           bounces, passes, steps, sec, samples / sec / 1e6, exchange_ms / steps, vt_group_exchange_bytes(g.group()));
    return 0;
}
