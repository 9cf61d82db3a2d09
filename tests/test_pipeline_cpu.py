"""Input pipeline: the oracle against the unmodified reference dataset classes' outputs."""
import numpy as np

from oracle import pipeline


def test_oracle_matches_reference_dataset_classes(golden):
    g = golden("pipeline")
    assert g["mean"].dtype == np.float32 and np.array_equal(g["mean"], pipeline.CITYSCAPES_MEAN_BGR)
    for i in range(3):
        x = pipeline.image_to_input(g["image%d" % i], g["mean"])
        ref = g["input%d" % i]
        assert x.dtype == ref.dtype == np.float32 and x.shape == ref.shape
        assert np.array_equal(x, ref)                      # bit-exact
    # fixture.MEAN_RGB (bench / model tests) is this mean, reversed and rounded to 3 decimals
    from oracle import fixture
    assert np.allclose(np.asarray(fixture.MEAN_RGB), g["mean"][::-1], atol=1e-3)


def test_batch_layout():
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (2, 6, 10, 3)).astype(np.uint8)
    x = pipeline.batch_to_input(imgs)
    assert x.shape == (2, 3, 6, 10)
    assert x[1, 0, 2, 3] == np.float32(imgs[1, 2, 3, 2]) - pipeline.CITYSCAPES_MEAN_BGR[2]     # R plane = BGR channel 2


# ----------------------------------------------------------------------------------------------------------------------
# The DEVICE SOURCE of esn_image_u8hwc_to_f32nchw, executed on the CPU: csrc/esn_input_kernel.cuh is free of CUDA headers,
# tests/cuda_cpu_shim.h maps every CUDA thread of a CTA to a pthread (__syncthreads = pthread barrier, __shared__ = one
# static instance), tests/input_kernel_host.cpp restates the entry point's launch arithmetic.  Checks the kernel's
# indexing, ragged tails, both staging branches (16-byte vector / byte copy), both store branches (float4 / guarded
# scalar), several tiles per CTA and barrier placement -- bit-exact against the oracle.  Not a substitute for the device
# run (tests/test_zz_pipeline_gpu.py); it is what can be known about the kernel without one.
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kernel_on_cpu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("ikh") / "input_kernel_host")
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "tests"),
                    "-I" + os.path.join(ROOT, "efficient-segmentation-networks_b200", "csrc"),
                    os.path.join(ROOT, "tests", "input_kernel_host.cpp"), "-o", exe], check=True)

    def run(img, reverse, in_off=0, out_off=0, grid_cap=8):
        n, h, w, _ = img.shape
        r = subprocess.run([exe] + [str(v) for v in (n, h, w, int(reverse), in_off, out_off, grid_cap)],
                           input=img.tobytes(), capture_output=True, timeout=300, check=True)
        return np.frombuffer(r.stdout, dtype=np.float32).reshape(n, 3, h, w)
    return run


@pytest.mark.parametrize("case", [
    # (n, h, w, reverse, input byte offset, output float offset, grid cap)
    (1, 1, 1, 1, 0, 0, 8),          # one pixel: byte staging, scalar stores
    (2, 17, 23, 1, 0, 0, 8),        # h*w % 4 != 0: ragged tail, image starts off 16-byte alignment
    (3, 32, 32, 1, 0, 0, 2),        # exactly one full tile per image, CTAs walk two tiles
    (2, 33, 64, 0, 0, 0, 3),        # full tile + partial tile, channel order kept
    (2, 40, 64, 1, 1, 0, 2),        # input not 16-byte aligned -> byte staging on full tiles
    (2, 40, 64, 1, 0, 1, 2),        # output not 16-byte aligned -> guarded scalar stores
    (5, 31, 33, 1, 3, 3, 4),        # everything ragged
    (1, 64, 128, 1, 0, 0, 100),     # grid == tiles
])
def test_device_source_on_cpu_is_bit_exact(kernel_on_cpu, case):
    n, h, w, reverse, in_off, out_off, cap = case
    rng = np.random.RandomState(n * 7919 + h * 31 + w)
    img = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    y = kernel_on_cpu(img, reverse, in_off, out_off, cap)
    ref = (pipeline.batch_to_input(img) if reverse
           else (img.astype(np.float32) - pipeline.CITYSCAPES_MEAN_BGR).transpose(0, 3, 1, 2))
    assert np.array_equal(y, ref)


# --------------------------------------------------------------------------- training-time augmentation
def test_resize_restatement_is_bit_exact_against_cv2():
    """oracle/pipeline.py restates cv2.resize (cv2 is a third-party dependency of the reference, absent from /root/reference):
    INTER_LINEAR on uint8 (11-bit fixed point) and INTER_NEAREST, every scale factor of the reference, odd sizes, 1 and 3
    channels -- bit-exact against the cv2 installed here."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for h, w in ((64, 128), (37, 53), (33, 70), (100, 31), (5, 7)):
        for ch in (3, 1):
            img = rng.randint(0, 256, (h, w, ch)).astype(np.uint8)
            img = img[:, :, 0] if ch == 1 else img
            for f in pipeline.SCALES:
                assert np.array_equal(pipeline.resize_linear_u8(img, f), cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR))
                assert np.array_equal(pipeline.resize_nearest_u8(img, f), cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_NEAREST))


def _golden_cases():
    import random
    g = np.load(os.path.join(ROOT, "tests", "golden", "augment.npz"))
    for k in range(int(g["n_cases"][0])):
        i, ch, cw, seed = [int(v) for v in g["case%d" % k]]
        random.seed(seed)
        np.random.seed(seed)
        params = pipeline.draw_train_params(g["image%d" % i].shape[:2], (ch, cw))
        yield g, k, g["image%d" % i], g["label%d" % i], (ch, cw), params


def test_train_augmentation_oracle_matches_reference_class():
    """oracle.pipeline.train_item + draw_train_params against the UNMODIFIED CityscapesDataSet (tools/make_golden_augment.py):
    18 seeded samples, every scale factor, padding in no / one / both directions, both mirror states -- bit-exact."""
    scales = set()
    for g, k, img, lab, crop, params in _golden_cases():
        x, y = pipeline.train_item(img, lab, *params, crop, g["mean"])
        assert np.array_equal(x, g["x%d" % k]) and np.array_equal(y, g["y%d" % k]), k
        scales.add(params[0])
    assert scales == set(pipeline.SCALES)


@pytest.fixture(scope="module")
def augment_on_cpu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("aug") / "augment_kernel_host")
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "tests"),
                    "-I" + os.path.join(ROOT, "efficient-segmentation-networks_b200", "csrc"),
                    os.path.join(ROOT, "tests", "augment_kernel_host.cpp"), "-o", exe], check=True)
    return exe


def test_augment_device_source_on_cpu_matches_reference_class(augment_on_cpu):
    """The kernel's DEVICE SOURCE (csrc/esn_augment_kernel.cuh) compiled with g++ reproduces the reference class's crops and
    label crops bit for bit on all 18 golden samples."""
    for g, k, img, lab, (ch, cw), (f, h_off, w_off, flip) in _golden_cases():
        h, w = img.shape[:2]
        rh, rw = pipeline.resized_size(h, w, f)
        args = [h, w, ch, cw, rh, rw, repr(1.0 / f), h_off, w_off, int(flip < 0), 1, 255] + [repr(float(v)) for v in g["mean"]]
        r = subprocess.run([augment_on_cpu] + [str(v) for v in args], input=img.tobytes() + lab.tobytes(), capture_output=True,
                           timeout=300, check=True)
        n = 3 * ch * cw
        x = np.frombuffer(r.stdout[:4 * n], dtype=np.float32).reshape(3, ch, cw)
        y = np.frombuffer(r.stdout[4 * n:], dtype=np.int64).reshape(ch, cw)
        assert np.array_equal(x, g["x%d" % k]), k
        assert np.array_equal(y.astype(np.float32), g["y%d" % k]), k
