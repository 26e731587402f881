"""The fused kernel decides each quantized coefficient from a fast butterfly DCT only when the
reference's own sum provably rounds to the same integer: |s_ref - g*T| <= kGamma * sum|p| with
kGamma = 1e-5 (derivation in DESIGN.md).  This CPU test measures the actual ratio with a C
emulation of the device butterfly (same operation order, fmaf) against the reference-order sum
over several input families; the analytic bound must hold with a wide margin."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_guard_band_margin(tmp_path):
    exe = str(tmp_path / "guard_band_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "guard_band_check.c"), "-lm"], check=True)
    cuh = open(os.path.join(ROOT, "jpeg_image_compression_b200", "csrc", "common.cuh")).read()

    def const(name):
        return float(re.search(name + r"\s*=\s*([0-9.eE+-]+)f", cuh).group(1))
    gamma, ga, gc, g0 = const("kGamma"), const("kGammaA"), const("kGammaC"), const("kGamma0")
    out = subprocess.run([exe, "150000", "11", str(ga), str(gc), str(g0)], check=True, capture_output=True, text=True).stdout
    worst, worst_refined = (float(v) for v in out.split())
    assert 0 < worst < gamma / 5, (worst, gamma)
    assert 0 < worst_refined < 0.5, worst_refined           # refined (mean-separated) bound: at least 2x margin


def test_guard_band_margin_tensor_core(tmp_path):
    """The tensor-core transform (K1, TC = true): reference LUT products in 22-bit fixed point, exact integer limb sums,
    t = fmaf(S0, 2048, S1).  Same check against the reference-order sum, with the TC constants of common.cuh; also the
    weight-sum error that enters the A-term of the refined bound must stay below the 32 u it was budgeted with."""
    exe = str(tmp_path / "guard_band_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "guard_band_check.c"), "-lm"], check=True)
    cuh = open(os.path.join(ROOT, "jpeg_image_compression_b200", "csrc", "common.cuh")).read()

    def const(name):
        return float(re.search(name + r"\s*=\s*([0-9.eE+-]+)f \* kTcScale", cuh).group(1))
    gamma, ga, gc, g0 = const("kTcGamma"), const("kTcGammaA"), const("kTcGammaC"), const("kTcGamma0")
    u = 2.0 ** -24
    assert gamma >= 76.2 * u and ga >= 15.4 * u and gc >= 74.2 * u and g0 >= 430 * u
    out = subprocess.run([exe, "150000", "11", str(ga), str(gc), str(g0), "tc"], check=True, capture_output=True, text=True).stdout
    worst, worst_refined, sum_err_u = (float(v) for v in out.split())
    assert 0 < worst < gamma / 5, (worst, gamma)
    assert 0 < worst_refined < 0.5, worst_refined
    assert sum_err_u <= 32.0, sum_err_u


def test_limb_matrix_matches_emulation():
    """The host code that builds the fp16 limb matrix (jpegb200.cu:build_tables) and the C emulation split the same
    fixed-point weights the same way."""
    cu = open(os.path.join(ROOT, "jpeg_image_compression_b200", "csrc", "jpegb200.cu")).read()
    emu = open(os.path.join(ROOT, "tests", "native", "guard_band_check.c")).read()
    for text in (cu, emu):
        assert "llround(w * 2097152.0)" in text
        assert "% 2048 + 2048) % 2048 - 1024" in text


def test_emulated_butterfly_matches_device_source():
    """The C emulation and the CUDA butterfly must be the same code (constants and op order)."""
    dev = open(os.path.join(ROOT, "jpeg_image_compression_b200", "csrc", "fused_block.cuh")).read()
    emu = open(os.path.join(ROOT, "tests", "native", "guard_band_check.c")).read()

    def body(text):
        m = re.search(r"x1 = fmaf\(b3, C7.*?x7 = fmaf\([^;]*;", text.replace("*x", "x"), re.S)
        return re.sub(r"\s+", "", m.group(0))
    assert body(dev) == body(emu)
    for const in ("0.98078528040323044913f", "0.83146961230254523708f", "0.55557023301960222474f",
                  "0.19509032201612826785f", "0.41421356237309504880f"):
        assert const in dev and const in emu
