// C++ surface test (include/open3d_b200.hpp): reads like the reference's own
// cpp/tests/t/pipelines/registration tests.  Exit code 0 = pass.
//   mode "host": argument validation + no-device behaviour (runs anywhere)
//   mode "gpu" : ICP + Model::Integrate on the device, checked against the CPU oracle
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "open3d_b200.hpp"
extern "C" {
#include "oracle.h"
}

namespace reg = open3d_b200::t::pipelines::registration;
namespace geo = open3d_b200::t::geometry;
namespace slam = open3d_b200::t::pipelines::slam;

#define EXPECT(c)                                                      \
    do {                                                               \
        if (!(c)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

template <class F>
static bool throws_with(F&& f, const char* needle) {
    try {
        f();
    } catch (const std::runtime_error& e) {
        return std::strstr(e.what(), needle) != nullptr;
    }
    return false;
}

static int host_mode() {
    float dummy[3] = {0, 0, 0};
    geo::PointCloud empty, src{dummy, nullptr, 1}, tgt_no_normals{dummy, nullptr, 1}, tgt{dummy, dummy, 1};
    EXPECT(throws_with([&] { reg::ICP(empty, tgt, 0.1); }, "empty"));
    EXPECT(throws_with([&] { reg::ICP(src, tgt_no_normals, 0.1); }, "normal"));
    EXPECT(throws_with([&] { reg::ICP(src, tgt, 0.0); }, "Max correspondence distance"));
    EXPECT(throws_with([&] { reg::ICP(src, tgt, 0.1, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}, reg::TransformationEstimationPointToPlane(), {}, 0.02); },
                       "VoxelDownSample"));
    // ColoredICP argument checks (Registration.cpp:160-176, TransformationEstimation.cpp:401-404)
    const std::array<double, 16> eye{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}};
    reg::TransformationEstimationForColoredICP colored;
    geo::PointCloud csrc{dummy, nullptr, 1, dummy}, ctgt{dummy, dummy, 1, dummy};
    EXPECT(throws_with([&] { reg::ICP(src, ctgt, 0.1, eye, colored); }, "source pointcloud to have colors"));
    EXPECT(throws_with([&] { reg::ICP(csrc, tgt, 0.1, eye, colored); }, "target pointcloud to have colors"));
    EXPECT(throws_with([&] { reg::ICP(csrc, ctgt, 0.1, eye, colored); }, "color_gradients"));
    EXPECT(reg::TransformationEstimationForColoredICP(7.0).lambda_geometric_ == 0.968 && colored.lambda_geometric_ == 0.968);
    reg::ICPConvergenceCriteria c;
    EXPECT(c.relative_fitness_ == 1e-6 && c.relative_rmse_ == 1e-6 && c.max_iteration_ == 30);
    reg::RobustKernel k;
    EXPECT(k.type_ == reg::RobustKernelMethod::L2Loss && k.scaling_parameter_ == 1.0 && k.shape_parameter_ == 1.0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        // no CPU fallback: the call must fail loudly
        EXPECT(throws_with([&] { reg::ICP(src, tgt, 0.1); }, "CUDA") || throws_with([&] { reg::ICP(src, tgt, 0.1); }, "cuda"));
        EXPECT(throws_with([&] { slam::Model m(0.008f); }, "CUDA") || throws_with([&] { slam::Model m(0.008f); }, "cuda"));
    }
    std::printf("cpp host-mode ok\n");
    return 0;
}

static int gpu_mode() {
    // a tilted plane with noise, moved by a small rigid motion
    const int side = 200, n = side * side;
    std::vector<float> tgt(3 * n), nrm(3 * n), src(3 * n);
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> jit(-0.4f, 0.4f);
    for (int i = 0; i < side; ++i)
        for (int j = 0; j < side; ++j) {
            const int k = i * side + j;
            const float x = (i + 0.5f + jit(rng)) * 0.02f, y = (j + 0.5f + jit(rng)) * 0.02f;
            const float z = 0.3f * std::sin(1.1f * x) * std::cos(0.9f * y) + 0.05f * std::sin(5 * x + 1);
            const float dzdx = 0.33f * std::cos(1.1f * x) * std::cos(0.9f * y) + 0.25f * std::cos(5 * x + 1);
            const float dzdy = -0.27f * std::sin(1.1f * x) * std::sin(0.9f * y);
            const float inv = 1.0f / std::sqrt(dzdx * dzdx + dzdy * dzdy + 1);
            tgt[3 * k] = x; tgt[3 * k + 1] = y; tgt[3 * k + 2] = z;
            nrm[3 * k] = -dzdx * inv; nrm[3 * k + 1] = -dzdy * inv; nrm[3 * k + 2] = inv;
            src[3 * k] = x + 0.012f; src[3 * k + 1] = y - 0.008f; src[3 * k + 2] = z + 0.01f;
        }
    float *d_src, *d_tgt, *d_nrm;
    int64_t* d_corr;
    EXPECT(cudaMalloc(&d_src, sizeof(float) * 3 * n) == cudaSuccess);
    EXPECT(cudaMalloc(&d_tgt, sizeof(float) * 3 * n) == cudaSuccess);
    EXPECT(cudaMalloc(&d_nrm, sizeof(float) * 3 * n) == cudaSuccess);
    EXPECT(cudaMalloc(&d_corr, sizeof(int64_t) * n) == cudaSuccess);
    cudaMemcpy(d_src, src.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tgt, tgt.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_nrm, nrm.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
    geo::PointCloud S{d_src, nullptr, n}, T{d_tgt, d_nrm, n};
    int calls = 0;
    auto res = reg::ICP(S, T, 0.05, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}, reg::TransformationEstimationPointToPlane(),
                        reg::ICPConvergenceCriteria(0, 0, 8), -1.0, [&](int, double, double) { ++calls; }, d_corr);
    EXPECT(calls == 8 && res.num_iterations_ == 8 && !res.converged_);
    const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    orc_icp_result ref;
    std::vector<double> per(16);
    EXPECT(orc_icp_p2plane_f32(src.data(), n, tgt.data(), nrm.data(), n, 0.05, I, 8, 0, 0, 0, 1.0, 1.0, 1, &ref, per.data(),
                               nullptr) == 0);
    for (int i = 0; i < 16; ++i) EXPECT(std::fabs(res.transformation_[i] - ref.transformation[i]) < 2e-5);
    EXPECT(std::fabs(res.fitness_ - ref.fitness) < 2e-4 && std::fabs(res.inlier_rmse_ - ref.inlier_rmse) < 2e-6);
    EXPECT(res.per_iteration_.size() == 8 && res.per_iteration_[0][0] == per[0]);
    // ColoredICP on the same pair: texture in the target frame, gradients from EstimateColorGradients
    {
        std::vector<float> sc(3 * n), tc(3 * n), grad(3 * n), grad_ref(3 * n);
        auto tex = [](float x, float y, int c) { return 0.5f + 0.5f * std::sin((2.1f + 0.7f * c) * x + (1.3f + c) * y + c); };
        for (int k = 0; k < n; ++k)
            for (int ch = 0; ch < 3; ++ch) {
                tc[3 * k + ch] = tex(tgt[3 * k], tgt[3 * k + 1], ch);
                sc[3 * k + ch] = tex(src[3 * k] - 0.012f, src[3 * k + 1] + 0.008f, ch);   // colour of where it belongs
            }
        float *d_sc, *d_tc, *d_grad;
        EXPECT(cudaMalloc(&d_sc, sizeof(float) * 3 * n) == cudaSuccess && cudaMalloc(&d_tc, sizeof(float) * 3 * n) == cudaSuccess &&
               cudaMalloc(&d_grad, sizeof(float) * 3 * n) == cudaSuccess);
        cudaMemcpy(d_sc, sc.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(d_tc, tc.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
        geo::PointCloud CS{d_src, nullptr, n, d_sc}, CT{d_tgt, d_nrm, n, d_tc};
        CT.EstimateColorGradients(d_grad, 30, 0.1);
        cudaMemcpy(grad.data(), d_grad, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost);
        orc_estimate_color_gradients_f32(tgt.data(), nrm.data(), tc.data(), n, 0.1, 30, grad_ref.data());
        double gmax = 0, gerr = 0;
        for (int k = 0; k < 3 * n; ++k) {
            gmax = std::fmax(gmax, std::fabs(grad_ref[k]));
            gerr = std::fmax(gerr, std::fabs(grad[k] - grad_ref[k]));
        }
        EXPECT(gmax > 0.1 && gerr <= 1e-5 * gmax);
        auto cres = reg::ICP(CS, CT, 0.05, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}},
                             reg::TransformationEstimationForColoredICP(0.968), reg::ICPConvergenceCriteria(0, 0, 6));
        orc_icp_result cref;
        EXPECT(orc_icp_colored_f32(src.data(), sc.data(), n, tgt.data(), nrm.data(), tc.data(), grad_ref.data(), n, 0.05, I, 6,
                                   0, 0, 0.968, 0, 1.0, 1.0, &cref, per.data(), nullptr) == 0);
        EXPECT(cres.num_iterations_ == 6);
        for (int i = 0; i < 16; ++i) EXPECT(std::fabs(cres.transformation_[i] - cref.transformation[i]) < 2e-5);
        cudaFree(d_sc);
        cudaFree(d_tc);
        cudaFree(d_grad);
    }
    // robust kernel + singular system error text
    EXPECT(throws_with(
            [&] {
                std::vector<float> z(3 * 64, 0.f), up(3 * 64, 0.f);
                for (int i = 0; i < 64; ++i) up[3 * i + 2] = 1.f;
                cudaMemcpy(d_src, z.data(), sizeof(float) * 3 * 64, cudaMemcpyHostToDevice);
                cudaMemcpy(d_tgt, z.data(), sizeof(float) * 3 * 64, cudaMemcpyHostToDevice);
                cudaMemcpy(d_nrm, up.data(), sizeof(float) * 3 * 64, cudaMemcpyHostToDevice);
                geo::PointCloud s2{d_src, nullptr, 64}, t2{d_tgt, d_nrm, 64};
                reg::ICP(s2, t2, 0.05);
            },
            "Singular 6x6"));
    // slam::Model on a constant-depth frame (a wall 1.5 m away)
    std::vector<uint16_t> depth(480 * 640, 1500);
    uint16_t* d_depth;
    EXPECT(cudaMalloc(&d_depth, depth.size() * 2) == cudaSuccess);
    cudaMemcpy(d_depth, depth.data(), depth.size() * 2, cudaMemcpyHostToDevice);
    slam::Model model(0.008f, 16, 4000);
    slam::Frame f;
    f.height = 480; f.width = 640; f.depth = d_depth;
    model.UpdateFramePose(0, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
    model.Integrate(f);
    std::vector<int32_t> keys(3 * 76800);
    const double K[9] = {525.0, 0, 319.5, 0, 525.0, 239.5, 0, 0, 1};
    const int64_t want = orc_depth_touch(depth.data(), 0, 480, 640, K, I, 16, 0.008f, 0.064f, 1000.0f, 3.0f, 4, keys.data(), 76800);
    EXPECT(want > 100 && model.NumBlocks() == want);
    f.depth = depth.data();
    f.images_on_host = true;
    model.Integrate(f);   // same frame again through the host-image path: no new blocks
    EXPECT(model.NumBlocks() == want);
    // Model::SynthesizeModelFrame: the wall comes back at 1.5 m (O(voxel) staircase of the march)
    {
        float* d_out;
        EXPECT(cudaMalloc(&d_out, sizeof(float) * 480 * 640) == cudaSuccess);
        // weight_threshold = min(frame_id, 3) (Model.cpp:45-47): at frame 0 it is 0 and never-observed voxels
        // (tsdf 0, weight 0) count as surface, as upstream; the SLAM loop only synthesizes from frame 1 on
        model.UpdateFramePose(1, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
        model.SynthesizeModelFrame(480, 640, f.intrinsics, d_out, nullptr);
        std::vector<float> out(480 * 640);
        cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
        int64_t hits = 0, close = 0;
        for (float d : out)
            if (d > 0) {
                ++hits;
                close += std::fabs(d - 1500.f) < 16.f;
            }
        EXPECT(hits > 0.9 * out.size() && close > 0.99 * hits);
        namespace odo = open3d_b200::t::pipelines::odometry;
        EXPECT(throws_with([&] { odo::RGBDOdometryMultiScale({d_depth, false, 480, 640}, {d_out, true, 480, 640}, f.intrinsics); },
                           "PointToPlane"));
        // Model::TrackFrameToModel on a curved surface (all 6 DoF observable; a plane would leave x / y / roll free):
        // the frame tracked against the model it was just fused into comes back as ~identity
        std::vector<uint16_t> curved(480 * 640);
        for (int v = 0; v < 480; ++v)
            for (int u = 0; u < 640; ++u)
                curved[v * 640 + u] = (uint16_t)std::lround(1500.0 + 220.0 * std::sin(u / 70.0) * std::cos(v / 55.0) + 0.25 * (v - 240));
        cudaMemcpy(d_depth, curved.data(), curved.size() * 2, cudaMemcpyHostToDevice);
        slam::Model m2(0.008f, 16, 6000);
        slam::Frame f2;
        f2.height = 480; f2.width = 640; f2.depth = d_depth;
        m2.UpdateFramePose(0, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
        m2.Integrate(f2);
        m2.UpdateFramePose(1, {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
        m2.Integrate(f2);
        m2.SynthesizeModelFrame(480, 640, f2.intrinsics, d_out, nullptr);
        const auto tr = m2.TrackFrameToModel(f2, d_out);
        EXPECT(tr.fitness_ > 0.8);
        for (int i = 0; i < 16; ++i) EXPECT(std::fabs(tr.transformation_[i] - (i % 5 == 0 ? 1.0 : 0.0)) < 0.02);
        cudaFree(d_out);
    }
    std::printf("cpp gpu-mode ok: fitness %.4f rmse %.5f blocks %lld\n", res.fitness_, res.inlier_rmse_, (long long)want);
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "host";
    try {
        return mode == "gpu" ? gpu_mode() : host_mode();
    } catch (const std::exception& e) {
        std::printf("unexpected exception: %s\n", e.what());
        return 2;
    }
}
