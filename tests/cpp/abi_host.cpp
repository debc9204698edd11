// A C++ host of the C ABI, built with plain g++ against include/ppg_b200.h and linked to libppg_b200.so -- what the
// reference's own C++ (System / Tracking / Matcher) does after the drop-in.  Everything here runs without a GPU:
// configuration defaults, the host-side vocabulary reader, the error path of ppg_create on a box without a device.
// On a GPU box the same binary goes on to extract one synthetic frame and run ExtendMapMatches through the shim-free
// C calls (argv[2] == "gpu").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ppg_b200.h"

#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                   \
        }                                                               \
    } while (0)

int main(int argc, char** argv) {
    CHECK(argc >= 2);
    const std::string weights_dir = argv[1];
    CHECK(ppg_api_version() == PPG_API_VERSION);
    ppg_config cfg;
    ppg_default_config(&cfg);
    CHECK(cfg.junction_max_num == 500 && cfg.junction_nms_radius == 4 && cfg.heatmap_refine_sz == 16);
    CHECK(std::fabs(cfg.junction_thresh - 1.0f / 128.0f) < 1e-9f && std::fabs(cfg.th_high - 0.8f) < 1e-7f);

    // the reference's vocabulary (exported blob) through the library's reader
    ppg_voc_file* vf = nullptr;
    ppg_vocabulary voc;
    const std::string vpath = weights_dir + "/voc_euroc_9x3.bin";
    CHECK(ppg_vocabulary_open(vpath.c_str(), &vf, &voc) == PPG_OK);
    CHECK(voc.k == 9 && voc.L == 3 && voc.n_nodes == 820 && voc.dim == PPG_DESC_DIM && voc.scoring == 1);
    int leaves = 0, words = 0;
    for (int i = 0; i < voc.n_nodes; i++) {
        leaves += voc.children[(size_t)i * voc.k] < 0;
        words += voc.word_id[i] >= 0;
    }
    CHECK(leaves == 729 && words == 729);
    ppg_vocabulary_close(vf);
    CHECK(ppg_vocabulary_open((weights_dir + "/does_not_exist.bin").c_str(), &vf, &voc) == PPG_ERR_WEIGHTS);
    CHECK(std::strlen(ppg_vocabulary_error()) > 0);

    // EuRoC camera (config/EuRoC.yaml:11-19 of the reference)
    cfg.width = 752;
    cfg.height = 480;
    const float K[9] = {458.654f, 0.f, 367.215f, 0.f, 457.296f, 248.375f, 0.f, 0.f, 1.f};
    const float D[4] = {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f};
    std::memcpy(cfg.K, K, sizeof(K));
    std::memcpy(cfg.D, D, sizeof(D));
    const std::string wpath = weights_dir + "/ppg_weights.bin";
    cfg.weights_path = wpath.c_str();
    ppg_ctx* ctx = nullptr;
    const int rc = ppg_create(&cfg, &ctx);
    if (argc < 3 || std::strcmp(argv[2], "gpu") != 0) {
        // no device here: the library must refuse, not fall back to the CPU
        CHECK(rc == PPG_ERR_CUDA && ctx == nullptr && std::strlen(ppg_last_error(nullptr)) > 0);
        std::puts("abi_host ok (cpu)");
        return 0;
    }
    CHECK(rc == PPG_OK && ctx != nullptr);
    std::vector<uint8_t> img((size_t)cfg.width * cfg.height, 90);
    for (int y = 100; y < 300; y++)
        for (int x = 200; x < 500; x++) img[(size_t)y * cfg.width + x] = (uint8_t)(((x / 40 + y / 40) & 1) ? 200 : 30);
    const uint8_t* frames[1] = {img.data()};
    ppg_frame_out out;
    CHECK(ppg_extract(ctx, frames, nullptr, 1, &out) == PPG_OK);
    CHECK(out.n_kp > 10 && out.n_kp <= 500 && out.status == 0);
    // a map made of the frame's own keypoints: every map point must find its keypoint again
    const int N = out.n_kp;
    std::vector<float> uv(2 * (size_t)N), vc(N, 0.95f);
    std::vector<uint8_t> cand(N, 1), obs(N, 1), bad(N, 0);
    std::vector<int32_t> eoff(N + 1, 0), kp_mp(N), kedge(out.n_edges > 0 ? out.n_edges : 1);
    for (int i = 0; i < N; i++) {
        uv[2 * i] = out.kp_x[i];
        uv[2 * i + 1] = out.kp_y[i];
    }
    CHECK(ppg_upload_map(ctx, out.desc, N) == PPG_OK);
    ppg_map_graph g{N, cand.data(), obs.data(), bad.data(), eoff.data(), nullptr, nullptr};
    CHECK(ppg_upload_map_graph(ctx, &g) == PPG_OK);
    std::vector<float> kx(out.kp_x, out.kp_x + N), ky(out.kp_y, out.kp_y + N), fd(out.desc, out.desc + (size_t)N * 256);
    std::vector<int32_t> es(out.edge_start, out.edge_start + out.n_edges), ee(out.edge_end, out.edge_end + out.n_edges);
    std::vector<int32_t> coff(out.conn_off, out.conn_off + N + 1), cidx(out.conn_idx, out.conn_idx + out.conn_off[N]);
    ppg_extend_in in{};
    in.n_kp = N;
    in.kp_x = kx.data();
    in.kp_y = ky.data();
    in.frame_desc = fd.data();
    in.n_edges = out.n_edges;
    in.edge_start = es.data();
    in.edge_end = ee.data();
    in.conn_off = coff.data();
    in.conn_idx = cidx.data();
    in.proj_uv = uv.data();
    in.view_cos = vc.data();
    in.th = 10.f;
    in.ratio = 0.8f;
    ppg_extend_out xo{};
    xo.kp_mp = kp_mp.data();
    xo.kedge_me = kedge.data();
    CHECK(ppg_extend_map_matches(ctx, &in, &xo) == PPG_OK);
    int self = 0;
    for (int i = 0; i < N; i++) self += kp_mp[i] == i;
    CHECK(xo.n_kp == N && self >= N - N / 10);  // distance 0 to itself; a few keypoints may be out of the grid
    // Matcher::SearchByProjection whole (ppg_search_by_projection): the same table projected onto the frame's own
    // keypoints, nothing assigned yet -- every row takes its own keypoint (distance 0 is the first minimum)
    std::vector<int32_t> kp_mp2(N, -7);
    ppg_projection_match_in pin{};
    pin.n_rows = N;
    pin.proj_uv = uv.data();
    pin.observed = nullptr;
    pin.n = N;
    pin.kp_x = kx.data();
    pin.kp_y = ky.data();
    pin.desc = fd.data();
    pin.kp_mp = nullptr;
    pin.th = 15.f;
    pin.max_dist = 0.8f;
    ppg_projection_match_out pout{};
    pout.kp_mp = kp_mp2.data();
    CHECK(ppg_search_by_projection(ctx, &pin, &pout) == PPG_OK);
    int self2 = 0;
    for (int i = 0; i < N; i++) self2 += kp_mp2[i] == i;
    CHECK(pout.nmatches >= self2 && pout.nmatches <= N && self2 >= N - N / 10);
    std::printf("abi_host ok (gpu): %d keypoints, %d edges, %d matched to themselves, nmatches %d; projection matcher %d\n", N,
                out.n_edges, self, xo.nmatches, pout.nmatches);
    ppg_destroy(ctx);
    return 0;
}
