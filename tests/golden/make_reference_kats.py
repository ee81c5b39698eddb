#!/usr/bin/env python
"""Extract the reference's own inline known-answer fixtures into reference_kats.json.

Run in the build container (needs /root/reference; the GPU box does not have
it, which is why the result is committed).  Nothing is computed here: every
number is parsed verbatim from Open3D's C++ test sources, and each entry records
the file:line span it came from.

    python tests/golden/make_reference_kats.py
"""
import json
import os
import re

REF = os.environ.get("O3D_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def line_of(text, pos):
    return text.count("\n", 0, pos) + 1


def init_block(text, var, start=0):
    """Parse `var = core::Tensor::Init<T>({...}, device)` -> nested list + line span."""
    m = re.search(re.escape(var) + r"\s*=\s*(?:core::)?Tensor::Init<(\w+)>\(", text[start:])
    assert m, var
    i = start + m.end()
    depth = 0
    j = i
    while True:
        c = text[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = text[i:j + 1]
    py = body.replace("{", "[").replace("}", "]")
    data = json.loads(py)
    return data, (line_of(text, start + m.start()), line_of(text, j)), j


def main():
    kats = {}

    # --- TransformationEstimation fixture ---------------------------------
    rel = "cpp/tests/t/pipelines/registration/TransformationEstimation.cpp"
    t = read(rel)
    src, l1, p = init_block(t, "core::Tensor source_points")
    tgt, l2, p = init_block(t, "core::Tensor target_points", p)
    nrm, l3, p = init_block(t, "core::Tensor target_normals", p)
    m = re.search(r"core::Tensor corres = core::Tensor::Init<int64_t>\(\s*\{([^}]*)\}", t)
    corr = [int(x) for x in m.group(1).split(",")]
    exp = {}
    for name, test in [("p2p_rmse", "ComputeRMSEPointToPoint"),
                       ("p2p_rmse_after", "ComputeTransformationPointToPoint"),
                       ("p2plane_rmse", "ComputeRMSEPointToPlane"),
                       ("p2plane_rmse_after", "ComputeTransformationPointToPlane")]:
        mt = re.search(r"TEST_P\(TransformationEstimationPermuteDevices,\s*" + test + r"\)", t)
        me = re.search(r"EXPECT_NEAR\(\w+,\s*(" + NUM + r"),\s*(" + NUM + r")\)", t[mt.end():])
        exp[name] = {"value": float(me.group(1)), "tol": float(me.group(2)),
                     "line": line_of(t, mt.end() + me.start())}
    kats["transformation_estimation"] = {
        "source": f"{rel}:{l1[0]}-{line_of(t, m.end())}",
        "source_points": src, "target_points": tgt, "target_normals": nrm,
        "correspondences": corr, "expected": exp}

    # --- NNS hybrid search ------------------------------------------------
    rel = "cpp/tests/core/NearestNeighborSearch.cpp"
    t = read(rel)
    s = t.index("TEST_P(NNSPermuteDevices, HybridSearch)")
    pts, l1, p = init_block(t, "core::Tensor dataset_points", s)
    mq = re.search(r"query_points\s*=\s*core::Tensor::Init<float>\(\{\{([^}]*)\}\}", t[s:])
    q = [float(x) for x in mq.group(1).split(",")]
    mr = re.search(r"double radius = (" + NUM + r");", t[s:])
    mk = re.search(r"int max_knn = (\d+);", t[s:])
    mi = re.search(r"gt_indices = core::Tensor::Init<int32_t>\(\{\{([^}]*)\}\}", t[s:])
    md = re.search(r"gt_distances =\s*core::Tensor::Init<float>\(\{\{([^}]*)\}\}", t[s:])
    mc = re.search(r"gt_counts = core::Tensor::Init<int32_t>\(\{(\d+)\}", t[s:])
    kats["hybrid_search"] = {
        "source": f"{rel}:{line_of(t, s)}-{line_of(t, s + mc.end())}",
        "dataset_points": pts, "query_points": [q], "radius": float(mr.group(1)),
        "max_knn": int(mk.group(1)),
        "gt_indices": [[int(x) for x in mi.group(1).split(",")]],
        "gt_distances": [[float(x) for x in md.group(1).split(",")]],
        "gt_counts": [int(mc.group(1))]}

    # --- robust kernel weights ----------------------------------------------
    rel = "cpp/tests/t/pipelines/registration/Registration.cpp"
    t = read(rel)
    s = t.index("TEST_P(RegistrationPermuteDevices, RobustKernel)")
    e = t.index("TEST_P(RegistrationPermuteDevices, GetInformationMatrixFromPointCloud)")
    body = t[s:e]
    table = re.findall(r"\{(\d),\s*(" + NUM + r")\}", body)
    gen = re.findall(r"scaling_parameter,\s*(" + NUM + r"),\s*\[&\]\(\)\s*\{\s*auto weight = "
                     r"GetWeightFromRobustKernel\((" + NUM + r")\);\s*EXPECT_NEAR\(weight,\s*("
                     + NUM + r"),\s*(" + NUM + r")\)", body)
    kats["robust_kernel"] = {
        "source": f"{rel}:{line_of(t, s)}-{line_of(t, e) - 1}",
        "scaling_parameter": 1.0, "shape_parameter": 1.0, "residual": 0.98, "tol": 1e-3,
        "expected_by_method": {k: float(v) for k, v in table},
        "generalized_by_shape": [{"shape": float(a), "residual": float(b), "expected": float(c),
                                  "tol": float(d)} for a, b, c, d in gen]}

    # --- VoxelBlockGrid indexing (set semantics of Activate) ---------------
    rel = "cpp/tests/t/geometry/VoxelBlockGrid.cpp"
    t = read(rel)
    s = t.index("TEST_P(VoxelBlockGridPermuteDevices, Indexing)")
    mk = re.search(r"std::vector<int>\{([^}]*)\}", t[s:])
    keys = [int(x) for x in mk.group(1).split(",")]
    mu = re.search(r"EXPECT_EQ\(buf_indices.GetLength\(\),\s*(\d+)\)", t[s:])
    kats["vbg_indexing"] = {
        "source": f"{rel}:{line_of(t, s)}-{line_of(t, s + mu.end())}",
        "keys": [keys[i:i + 3] for i in range(0, len(keys), 3)],
        "expected_unique": int(mu.group(1))}

    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
