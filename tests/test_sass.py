"""Static checks on the built sm_100a code (no GPU needed: cuobjdump reads the in-tree library).

* the tensor-core kernels really are tcgen05 / TMA code (UTCHMMA, UTMALDG, UTCBAR in the SASS);
* the single-thread issue loops stay lean: with `if (lane == 0)` around the producer / MMA roles nvcc wrapped every
  tcgen05 / TMA instruction in an `ELECT ... R2UR ... BRA.U.ANY` waterfall loop (125 instructions per K block — the
  tensor pipe of the conv kernels sat at 77 % with the issuing thread as the limit). Under
  `elect.sync` those loops do not exist; this test keeps it that way."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "instantrestore_b200" / "libinstantrestore_b200.so"


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    if not LIB.exists():
        from instantrestore_b200.build import build_library
        build_library()
    out = subprocess.run([exe, "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    return funcs


def _kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_tensor_core_kernels_are_tcgen05_and_tma(sass):
    for needle in ("conv_gemm_kernel", "conv_gemm_persistent_kernel", "conv_gemm_pair_kernel", "conv3_halo_kernel", "shared_attn_kernel"):
        ks = _kernels(sass, needle)
        assert ks, needle
        for name, lines in ks.items():
            text = "\n".join(lines)
            assert "UTCHMMA" in text, f"{name}: no tcgen05.mma"
            assert "UTMALDG" in text, f"{name}: no TMA load"
            assert "UTCBAR" in text, f"{name}: no tcgen05.commit"
            assert not re.search(r"(?<!UTC)HMMA\.", text), f"{name}: legacy mma.sync found"
    pair = "\n".join("\n".join(v) for v in _kernels(sass, "conv_gemm_pair_kernel").values())
    assert "UTCHMMA.2CTA" in pair and "UTMALDG.4D.2CTA" in pair and "UTCBAR.2CTA.MULTICAST" in pair


def test_issue_loops_have_no_uniform_register_waterfalls(sass):
    for needle in ("conv_gemm_kernel", "conv_gemm_persistent_kernel", "conv_gemm_pair_kernel", "conv3_halo_kernel", "shared_attn_kernel"):
        for name, lines in _kernels(sass, needle).items():
            n = sum("BRA.U.ANY" in l for l in lines)
            assert n == 0, f"{name}: {n} R2UR waterfall loops around tcgen05/TMA instructions (use elect.sync, not lane == 0)"


def test_mma_issue_loop_is_short(sass):
    """Between the full-barrier wait and the commit of a K block the CTA-pair kernel issues four MMAs; the whole loop body
    must stay far below the 512 clk of tensor work it feeds (it was 125 instructions; it is ~50)."""
    for name, lines in _kernels(sass, "conv_gemm_pair_kernelILi256").items():
        idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
        assert len(idx) >= 4
        code = [l for l in lines[idx[0]:idx[3] + 1] if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
        assert len(code) <= 16, f"{name}: {len(code)} instructions between the first and the fourth MMA of a K block"


def test_tma_store_epilogue_is_a_bulk_tensor_store(sass):
    """The TSTORE instantiation of the persistent kernel leaves through UTMASTG (cp.async.bulk.tensor ... global.shared::cta)
    and keeps no per-thread global store in its main epilogue path other than the statistics."""
    found = {n: l for n, l in _kernels(sass, "conv_gemm_persistent_kernel").items() if n.endswith("ELb0ELb1EEEvNS_11GemmKParamsE")}
    assert found, "no TSTORE instantiation of conv_gemm_persistent_kernel in the library"
    for name, lines in found.items():
        text = "\n".join(lines)
        assert "UTMASTG.2D" in text, f"{name}: no TMA store"


def test_halo_pair128_epilogue_runs_on_tensor_maps(sass):
    """The 128-wide CTA-pair halo kernel moves its residual in and its outputs out with TMA (UTMALDG / UTMASTG) and keeps
    no vectorised per-thread global store of output rows (STG.E.ENL2.256 / STG.E.128); the statistics' 8-byte stores stay."""
    found = _kernels(sass, "conv3_halo_kernelILi128ELb1E")
    assert len(found) == 2, "expected the residual and the no-residual instantiation of conv3_halo_kernel<128, pair>"
    for name, lines in found.items():
        text = "\n".join(lines)
        assert "UTMASTG.2D" in text, f"{name}: no TMA store"
        assert not re.search(r"STG\.E(\.ENL2)?\.(128|256)", text), f"{name}: per-thread row stores left in the epilogue"
        assert "UTCHMMA.2CTA" in text, f"{name}: not a cta_group::2 kernel"
