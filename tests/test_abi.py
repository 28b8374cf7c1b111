"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports exactly the entry
points include/oatgpu.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

import oat_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "oatgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(oat_[a-z0-9_]+)\s*\(", hdr))


def test_library_is_built_and_loads():
    assert os.path.exists(oat_b200.LIB_PATH), "run __graft_entry__.build() first"
    assert oat_b200.lib().oat_abi_version() == 1


def test_exports_every_declared_symbol():
    L = C.CDLL(oat_b200.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/oatgpu.h but not exported"
    assert declared == set(oat_b200.EXPORTED_SYMBOLS), declared ^ set(oat_b200.EXPORTED_SYMBOLS)


def test_only_the_abi_is_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", oat_b200.LIB_PATH], capture_output=True, text=True).stdout
    syms = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    assert syms == header_symbols(), syms ^ header_symbols()


def test_default_params_match_reference_defaults():
    """MOG2 defaults and HSVDetector defaults (HSVDetector.h:86-94, .cpp:42-43)."""
    p = oat_b200.default_mog_params()
    assert (p.history, p.nmixtures, p.var_threshold, p.var_threshold_gen) == (500, 5, 16.0, 9.0)
    assert abs(p.background_ratio - 0.9) < 1e-7 and (p.var_init, p.var_min, p.var_max) == (15.0, 4.0, 75.0)
    assert abs(p.complexity_reduction_threshold - 0.05) < 1e-8
    assert (p.detect_shadows, p.shadow_value, p.shadow_threshold) == (1, 127, 0.5)
    h = oat_b200.HsvParams()
    oat_b200.lib().oat_hsv_default_params(C.byref(h))
    assert (h.h_min, h.h_max, h.s_min, h.s_max, h.v_min, h.v_max) == (0, 256, 0, 256, 0, 256)
    assert (h.erode_px, h.dilate_px, h.min_area, h.max_area) == (0, 10, 0.0, oat_b200.DBL_MAX)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail with an error, never compute."""
    if oat_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(oat_b200.OatError) as e:
        oat_b200.Context(0)
    assert e.value.code == -2 and "CUDA" in str(e.value)
    h = C.c_void_p()
    assert oat_b200.lib().oat_mog_create(None, 10, 10, None, C.byref(h)) != 0
    assert oat_b200.lib().oat_tracker_create(None, 10, 10, None, 0, C.byref(h)) != 0
    assert oat_b200.lib().oat_hsvdet_create(None, 10, 10, C.byref(h)) != 0
    assert b"null context" in oat_b200.lib().oat_last_error()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under oat_b200/ may reference it."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "oat_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f"{fn} imports oracle"
                assert "oat_oracle" not in src and "orc_" not in src, f"{fn} references the oracle"


def _simulate_draws(ntiles, grid, stages):
    """The producer lane's scheduler loop of mog_pipe_kernel (csrc/mog_pipe.cuh: next_tile/refill), transcribed:
    every CTA walks its static tiles first + seq*grid, draws once when it starts the last static one, and then
    once more after every valid number; numbers are handed out in arrival order (here: round-robin over the CTAs
    that are still drawing -- the COUNT does not depend on the order).  Returns (draws, tiles processed)."""
    counter = 0
    tiles = []
    state = []  # per CTA: [seq, ahead, ended]
    for b in range(grid):
        seq, ahead, ended = 0, None, False
        for k in range(stages):  # the initial refills
            if ended:
                break
            t = b + seq * grid
            if seq == stages - 1:
                ahead = counter + stages * grid
                counter += 1
            seq += 1
            if t >= ntiles:
                ended = True
            else:
                tiles.append(t)
        state.append([seq, ahead, ended])
    live = [s for s in state if not s[2]]
    while live:
        nxt = []
        for s in live:
            t = s[1]
            if t is None:          # never reached its last static tile: no number in hand -> nothing more to do
                s[2] = True
                continue
            if t < ntiles:
                tiles.append(t)
                s[1] = counter + stages * grid
                counter += 1
                nxt.append(s)
            else:
                s[2] = True
        live = nxt
    return counter, tiles


def test_tile_scheduler_draw_count_matches_kernel_logic():
    """The host keeps the scheduler counters monotonic by knowing how many numbers each launch draws
    (api.cu: pipe_draws).  Pin that formula against the kernel's loop for every frame size / grid shape class,
    and check the loop hands out every tile exactly once."""
    L = oat_b200.lib()
    st = C.c_int()
    L.oat_debug_pipe_draws(1, 1, C.byref(st))
    S = st.value
    assert S >= 2
    cases = [(nt, g) for g in (1, 2, 3, 7, 148, 296, 592, 740) for nt in
             (g, g + 1, 2 * g - 1, 2 * g, 2 * g + 1, S * g - 1, S * g, S * g + 1, S * g + 37, 7 * g, 2025, 4050, 8100, 16200) if nt >= g]
    for ntiles, grid in cases:
        draws, tiles = _simulate_draws(ntiles, grid, S)
        assert sorted(tiles) == list(range(ntiles)), (ntiles, grid)
        assert L.oat_debug_pipe_draws(ntiles, grid, None) == draws, (ntiles, grid, draws)
