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


class _ResidentModel:
    """A transcription of the resident fused kernel's scheduling protocol (oat_b200/csrc/mog_pipe.cuh,
    mog_stream_kernel): every CTA has a LOADER lane, compute warps and `storers` STORER lanes around a ring of `stages`
    stages (loader --full--> compute --done--> storer --freed--> loader); work items (frame, tile) are drawn from
    one counter (the first `stages` of a CTA in one draw, then one at a time); a tile of frame f+1 of a model may
    only be loaded once the same tile of the model's previous frame has been PUBLISHED.  Storer lane x owns the stage
    uses x, x + storers, ...; it frees the stage first and publishes afterwards -- one lane: in batches (one release
    fence per batch), when a batch is full or, if the next tile does not arrive in time, while it waits; several lanes:
    every tile on its own, before the lane looks at its next one.  The end of the queue is marked in `storers`
    consecutive stage uses, one marker per lane, written together.
    The model is stepped by an adversarial (random) scheduler; it checks that every item is processed exactly
    once, that no load ever precedes the publication it depends on, and that the protocol cannot deadlock --
    including frames smaller than one CTA's ring, where the loader waits for a tile that still sits in its own
    CTA's stages."""

    def __init__(self, nframes, ntiles, grid, stages, models, rng, batch=2, storers=1):
        self.nf, self.nt, self.grid, self.S, self.rng, self.P, self.NS = nframes, ntiles, grid, stages, rng, batch, storers
        assert storers <= stages
        self.total = nframes * ntiles
        self.model_of = [f % models for f in range(nframes)]      # frames of `models` streams interleaved
        self.prev = {}                                            # frame -> previous frame of the same model
        last = {}
        for f in range(nframes):
            self.prev[f] = last.get(self.model_of[f])
            last[self.model_of[f]] = f
        self.counter = 0
        self.published = set()
        self.loaded = []
        self.exited = 0
        self.ctas = [self._new_cta() for _ in range(grid)]

    def _new_cta(self):
        # stage state: None (free) -> item (loaded) -> computed flag -> stored (free again)
        return {"stage": [None] * self.S, "li": 0, "si": list(range(self.NS)), "ci": 0, "pend": [[] for _ in range(self.NS)],
                "nxt": None, "batch": None, "ldone": False, "cdone": False, "sdone": [False] * self.NS}

    def _item(self, g):
        return "END" if g >= self.total else (g // self.nt, g % self.nt)

    def _ready(self, it):
        p = self.prev[it[0]]
        return p is None or (p, it[1]) in self.published

    def _draw(self, c):
        if c["batch"]:
            return c["batch"].pop(0)
        g = self.counter
        self.counter += 1
        return g

    def _freed(self, c, use):
        """mbar_wait(freed[use % S]) for the stage's previous tenant: use - S has been written back by its storer lane"""
        prev = use - self.S
        return prev < 0 or c["si"][prev % self.NS] > prev

    def step_loader(self, c):
        S = self.S
        if c["ldone"]:
            return False
        if c["batch"] is None:  # the first draw: `stages` consecutive items
            c["batch"] = [self.counter + k for k in range(1, S)]
            c["nxt"] = self._item(self.counter)
            self.counter += S
            return True
        if not self._freed(c, c["li"]):
            return False  # only this CTA's own storer lanes are waited for
        if c["nxt"] == "END":
            if not all(self._freed(c, c["li"] + k) for k in range(1, self.NS)):
                return False
            for k in range(self.NS):  # one marker per storer lane, all written before the compute warps are released
                assert c["stage"][(c["li"] + k) % S] is None
                c["stage"][(c["li"] + k) % S] = "END"
            c["li"] += self.NS
            c["ldone"] = True
            return True
        if not self._ready(c["nxt"]):
            return False  # spins on the tile's flag
        it = c["nxt"]
        self.loaded.append(it)
        assert c["stage"][c["li"] % S] is None
        c["stage"][c["li"] % S] = it
        c["nxt"] = self._item(self._draw(c))
        c["li"] += 1
        return True

    def step_compute(self, c):
        if c["cdone"] or c["ci"] >= c["li"]:
            return False
        s = c["ci"] % self.S
        if c["stage"][s] == "END":
            c["cdone"] = True
            c["ci"] += self.NS  # passes every lane's marker on
            return True
        c["ci"] += 1
        return True

    def step_storer(self, c, x=None):
        if x is None:
            x = self.rng.randrange(self.NS)
        if c["sdone"][x]:
            return False
        pend = c["pend"][x]
        if self.NS > 1 and pend:  # several lanes: the tile just written back is published before anything else
            self.published.update(pend)
            pend.clear()
            return True
        si = c["si"][x]
        s = si % self.S
        if si >= c["ci"]:  # done[s] has not completed: (one lane) after the bounded wait, publish what is owed
            if pend:
                self.published.update(pend)
                pend.clear()
                return True
            return False
        if c["stage"][s] == "END":
            self.published.update(pend)
            pend.clear()
            c["sdone"][x] = True
            self.exited += 1
            return True
        if len(pend) == self.P:  # a full batch: one fence publishes it
            self.published.update(pend)
            pend.clear()
        pend.append(c["stage"][s])
        c["stage"][s] = None
        c["si"][x] = si + self.NS
        return True

    def run(self, max_steps=10_000_000):
        steps = 0
        while self.exited < self.grid * self.NS:
            order = list(range(self.grid))
            self.rng.shuffle(order)
            progressed = False
            for b in order:
                c = self.ctas[b]
                acts = [self.step_loader, self.step_compute] + [lambda cc, x=x: self.step_storer(cc, x) for x in range(self.NS)]
                self.rng.shuffle(acts)
                for act in acts:
                    if self.rng.random() < 0.7:  # an adversarial scheduler: some actors simply do not run this round
                        progressed |= act(c)
            if not progressed:
                # nobody moved in a randomised round: give everybody a deterministic chance before calling it a deadlock
                for c in self.ctas:
                    progressed |= self.step_loader(c) | self.step_compute(c)
                    for x in range(self.NS):
                        progressed |= self.step_storer(c, x)
                assert progressed, "deadlock: no CTA can make progress"
            steps += 1
            assert steps < max_steps
        return self


def test_resident_scheduler_protocol_model():
    """Pins the resident kernel's work distribution + publication protocol on the CPU (see _ResidentModel)."""
    import random

    rng = random.Random(1234)
    cases = [(1, 1, 1, 1), (1, 7, 3, 1), (4, 5, 8, 1), (6, 2, 8, 2), (8, 13, 5, 1), (16, 3, 7, 4), (5, 40, 6, 1),
             (12, 9, 4, 3), (3, 1, 5, 1), (20, 2, 3, 2), (9, 1, 1, 1), (30, 2, 1, 1), (10, 4, 2, 2), (64, 75, 16, 1)]
    for nframes, ntiles, grid, models in cases:
        for S, P, NS in ((3, 1, 1), (4, 2, 1), (2, 3, 1), (4, 4, 1), (4, 1, 2), (3, 1, 2), (2, 1, 2), (4, 1, 3)):  # (shipped: 4 stages, 2 lanes)
            m = _ResidentModel(nframes, ntiles, grid, S, models, rng, batch=P, storers=NS).run()
            assert sorted(m.loaded) == [(f, t) for f in range(nframes) for t in range(ntiles)], (nframes, ntiles, grid)
            assert len(m.published) == nframes * ntiles


def test_doubling_window_model():
    """The band morphology's horizontal pass (oat_b200/csrc/tail_fast.cuh: hwin_or) ORs a k-wide window by doubling on the
    96 bits prev:cur:next of a word instead of one shift per window position.  A transcription of it against the
    definition -- bit x of the result = OR over [x - k/2, x - k/2 + k - 1] (cv::dilate's anchor, MORPH_RECT) -- for every
    k the fast path takes (1..32), random sparse and dense words; erosion is the same on the complement."""
    import random

    M64 = (1 << 64) - 1

    def hwin_or(prev, cur, nxt, k):
        lo, hi = (prev | (cur << 32)) & M64, nxt

        def shr_or(sft):
            nonlocal lo, hi
            nlo, nhi = ((lo >> sft) | (hi << (64 - sft))) & M64, hi >> sft
            lo |= nlo
            hi |= nhi

        w = 1
        while 2 * w <= k:
            shr_or(w)
            w <<= 1
        if w < k:
            shr_or(k - w)
        sft = 32 - k // 2
        return ((lo >> sft) | (hi << (64 - sft))) & 0xFFFFFFFF

    def window(prev, cur, nxt, k):
        P = prev | (cur << 32) | (nxt << 64)
        out = 0
        for x in range(32):
            if any((P >> (32 + x + d)) & 1 for d in range(-(k // 2), k - k // 2)):
                out |= 1 << x
        return out

    rng = random.Random(7)
    for k in range(1, 33):
        for density in (1, 2, 4):
            for _ in range(60):
                p, c, n = (rng.getrandbits(32) & rng.getrandbits(32) if density > 1 else rng.getrandbits(32) for _ in range(3))
                if density == 4:
                    p, c, n = p & rng.getrandbits(32), c & rng.getrandbits(32), n & rng.getrandbits(32)
                assert hwin_or(p, c, n, k) == window(p, c, n, k), (k, p, c, n)
                # erosion: complement in, complement out (all three words inside the image)
                e = ~hwin_or(~p & 0xFFFFFFFF, ~c & 0xFFFFFFFF, ~n & 0xFFFFFFFF, k) & 0xFFFFFFFF
                P = p | (c << 32) | (n << 64)
                want = 0
                for x in range(32):
                    if all((P >> (32 + x + d)) & 1 for d in range(-(k // 2), k - k // 2)):
                        want |= 1 << x
                assert e == want, (k, p, c, n)
