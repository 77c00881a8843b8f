"""Discrete-event model of the mbarrier protocols of the attention kernels (no GPU needed).

Each kernel's warp roles (TMA producer, MMA issuer, softmax / worker warps, dQ drain warps) are transcribed as Python
generators that wait on / arrive at model mbarriers, issue MMAs into an in-order tensor pipe (``tcgen05.commit`` = a marker
in that pipe) and touch model resources (smem stages, TMEM regions).  Random latencies are drawn per run.  The model checks:
  * no deadlock (every role terminates), no over-arrival on a barrier;
  * every wait names the phase it is waiting for, and the barrier is at most one phase ahead of it (no parity aliasing);
  * data hazards: a TMEM / smem region is only overwritten after all its readers are done with the previous contents, and
    only read when it holds the tile the reader expects.
It is a transcription, not the CUDA source: it guards the PROTOCOL (who waits for what, in which order), not layouts.
``python tools/sim_attn_protocol.py`` runs every kernel over sequence lengths 1..8 tiles with 200 random schedules each;
tests/test_attn_protocol.py runs a smaller sweep and two negative controls (a removed wait must be caught).
"""
from __future__ import annotations

import heapq
import random


class ProtocolError(AssertionError):
    pass


class Barrier:
    def __init__(self, sim, name, count):
        self.sim, self.name, self.count = sim, name, count
        self.arrived, self.phase, self.waiters = 0, 0, []

    def arrive(self, n=1):
        self.arrived += n
        if self.arrived > self.count:
            raise ProtocolError(f"{self.name}: {self.arrived} arrivals in a phase of {self.count}")
        if self.arrived == self.count:
            self.arrived = 0
            self.phase += 1
            ws, self.waiters = self.waiters, []
            for agent, k in ws:
                self.sim.ready(agent)


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.t = 0.0
        self.q = []          # (time, seq, fn)
        self.seq = 0
        self.agents = {}
        self.blocked = {}
        self.pipe = []       # in-order tensor pipe: list of ("mma", dur, on_start, on_end) / ("commit", barrier)
        self.pipe_busy = False

    def barrier(self, name, count):
        return Barrier(self, name, count)

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.q, (self.t + dt, self.seq, fn))

    def spawn(self, name, gen):
        self.agents[name] = gen
        self.ready(name)

    def ready(self, name):
        self.blocked.pop(name, None)
        self.at(self.rng.uniform(1, 40), lambda: self.step(name))

    def step(self, name):
        gen = self.agents.get(name)
        if gen is None:
            return
        try:
            req = next(gen)
        except StopIteration:
            del self.agents[name]
            return
        kind = req[0]
        if kind == "wait":          # ("wait", barrier, k): wait for completion of phase k
            _, bar, k = req
            if bar.phase not in (k, k + 1):
                raise ProtocolError(f"{name} waits for phase {k} of {bar.name} but the barrier is at phase {bar.phase}")
            if bar.phase > k:
                self.ready(name)
            else:
                bar.waiters.append((name, k))
                self.blocked[name] = (bar.name, k)
        elif kind == "sleep":
            self.at(req[1], lambda: self.step(name))
        elif kind == "now":         # zero-time action already performed by the generator
            self.at(self.rng.uniform(1, 10), lambda: self.step(name))
        else:
            raise ValueError(kind)

    # tensor pipe ------------------------------------------------------------------------------------------------
    def mma(self, dur, on_start=None, on_end=None):
        self.pipe.append(("mma", dur, on_start, on_end))
        self._pump()

    def commit(self, bar):
        self.pipe.append(("commit", bar))
        self._pump()

    def _pump(self):
        if self.pipe_busy or not self.pipe:
            return
        item = self.pipe.pop(0)
        if item[0] == "commit":
            lat = self.rng.uniform(20, 200)   # completion -> mbarrier arrival latency
            bar = item[1]
            self.at(lat, lambda: bar.arrive(1))
            self._pump()
            return
        _, dur, on_start, on_end = item
        self.pipe_busy = True
        if on_start:
            on_start()

        def done():
            if on_end:
                on_end()
            self.pipe_busy = False
            self._pump()
        self.at(dur * self.rng.uniform(0.9, 1.3), done)

    def tma(self, bar, on_land=None):
        def land():
            if on_land:
                on_land()
            bar.arrive(1)
        self.at(self.rng.uniform(800, 3000), land)

    def run(self):
        while self.q:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
        if self.agents:
            raise ProtocolError(f"deadlock: {sorted(self.agents)} never finished; blocked on {self.blocked}")


def check(cond, msg):
    if not cond:
        raise ProtocolError(msg)


# ================================================================================================ forward, two threads per row (fwd4)
def run_fwd4(nk, seed, bug=None):
    """attn_fwd4_kernel: nk = number of 128-key tiles; 8 softmax warps (two threads per row)."""
    sim = Sim(seed)
    B = lambda n, c: sim.barrier(n, c)
    q_full, s_full, s_free, p_ready, pv_done, o_full = B("q_full", 1), B("s_full", 1), B("s_free", 256), B("p_ready", 256), B("pv_done", 1), B("o_full", 1)
    k_full, k_empty = [B("k_full0", 1), B("k_full1", 1)], [B("k_empty0", 1), B("k_empty1", 1)]
    v_full, v_empty = [B("v_full0", 1), B("v_full1", 1)], [B("v_empty0", 1), B("v_empty1", 1)]
    S = dict(tile=-1, loaded=set())
    P = dict(tile=-1, written=set(), consumed=-1)
    O = dict(pv_running=False, done=-1)
    Kst, Vst = [dict(tile=-1, busy=False) for _ in range(2)], [dict(tile=-1, busy=False) for _ in range(2)]
    xch = dict(n=0)  # named barrier of the 8 softmax warps

    def tma():
        sim.tma(q_full)
        Kst[0]["tile"] = 0; sim.tma(k_full[0])
        Vst[0]["tile"] = 0; sim.tma(v_full[0])
        for j in range(1, nk):
            st = j & 1
            yield ("wait", k_empty[st], (j >> 1) - 1) if j >= 2 else ("now",)
            check(not Kst[st]["busy"], "K stage reloaded while in use")
            Kst[st]["tile"] = j; sim.tma(k_full[st])
            yield ("wait", v_empty[st], (j >> 1) - 1) if j >= 2 else ("now",)
            check(not Vst[st]["busy"], "V stage reloaded while in use")
            Vst[st]["tile"] = j; sim.tma(v_full[st])

    def issue_s(j):
        st = j & 1
        def start():
            check(Kst[st]["tile"] == j, f"S({j}) reads K stage holding {Kst[st]['tile']}")
            check(j == 0 or (S["tile"] == j - 1 and len(S["loaded"]) == 8), f"S({j}) overwrites scores {S['tile']} loaded by {len(S['loaded'])}/8 warps")
            Kst[st]["busy"] = True
        def end():
            S["tile"], S["loaded"] = j, set(); Kst[st]["busy"] = False
        sim.mma(256, start, end)
        sim.commit(s_full)
        sim.commit(k_empty[st])

    def mma():
        yield ("wait", q_full, 0)
        yield ("wait", k_full[0], 0)
        issue_s(0)
        for j in range(nk):
            st = j & 1
            if j + 1 < nk:
                if bug != "no_s_free":
                    yield ("wait", s_free, j)
                yield ("wait", k_full[st ^ 1], (j + 1) >> 1)
                issue_s(j + 1)
            yield ("wait", v_full[st], j >> 1)
            yield ("wait", p_ready, j)
            def start(j=j, st=st):
                check(P["tile"] == j and len(P["written"]) == 8, f"PV({j}) reads P holding {P['tile']} from {len(P['written'])}/8 warps")
                check(Vst[st]["tile"] == j, "PV reads the wrong V stage")
                O["pv_running"] = True; Vst[st]["busy"] = True
            def end(j=j, st=st):
                O["pv_running"] = False; O["done"] = j; P["consumed"] = j; Vst[st]["busy"] = False
            sim.mma(256, start, end)
            sim.commit(v_empty[st])
            sim.commit(pv_done)
        sim.commit(o_full)

    def softmax(w):
        for j in range(nk):
            yield ("wait", s_full, j)
            check(S["tile"] == j, f"softmax {w} loads scores of tile {S['tile']} expecting {j}")
            yield ("sleep", sim.rng.uniform(20, 500))
            check(S["tile"] == j, f"scores of tile {j} overwritten under softmax {w}'s load")
            S["loaded"].add(w)
            s_free.arrive(32)
            yield ("sleep", sim.rng.uniform(100, 400))    # maximum; then the smem exchange + named barrier (not modelled as a hazard)
            yield ("sleep", sim.rng.uniform(300, 1200))   # first 32 exponentials
            if j > 0 and bug != "no_pv_done":
                yield ("wait", pv_done, j - 1)
            check(P["consumed"] >= j - 1, f"P overwritten before PV({j - 1}) finished")
            if P["tile"] != j:
                P["tile"], P["written"] = j, set()
            yield ("sleep", sim.rng.uniform(300, 1200))   # second 32 exponentials
            if j > 0:
                check(O["done"] >= j - 1 and not O["pv_running"], "O rescaled while PV may run")
            P["written"].add(w)
            p_ready.arrive(32)
        yield ("wait", o_full, 0)

    sim.spawn("tma", tma()); sim.spawn("mma", mma())
    for w in range(8):
        sim.spawn(f"softmax{w}", softmax(w))
    sim.run()


# ================================================================================================ forward, persistent (fwd5)
# ================================================================================================ backward (bwd4)
def run_bwd4(nq, seed, bug=None):
    """attn_bwd4_kernel: nq query tiles; 16 worker warps (X then Y per tile), 4 dQ drain warps, 3-stage Q/dO ring, P^T / dS^T in
    tensor memory over the score columns, the per-query statistics in an auxiliary MMA operand tile rewritten one tile ahead."""
    sim = Sim(seed)
    B = lambda n, c: sim.barrier(n, c)
    NS, NW = 3, 16
    kv_full, s_full, dp_full, aux_init = B("kv_full", 1), B("s_full", 1), B("dp_full", 1), B("aux_init", 32 * NW)
    p_ready, ds_ready = B("p_ready", 32 * NW), B("ds_ready", 32 * NW)
    dq_full, dq_free, acc_full = B("dq_full", 1), B("dq_free", 128), B("acc_full", 1)
    q_full, q_empty = [B(f"q_full{k}", 1) for k in range(NS)], [B(f"q_empty{k}", 1) for k in range(NS)]
    Sreg, DPreg = dict(tile=-1, loaded=set()), dict(tile=-1, loaded=set())
    Pbuf = dict(tile=-1, written=set(), consumed=-1)        # P^T in TMEM over the score columns
    DSbuf = dict(tile=-1, written=set(), dk_done=-1, dq_done=-1)
    DQ = dict(tile=-1, drained=-1)
    Qst = [dict(tile=-1, busy=0) for _ in range(NS)]
    # auxiliary operand tile: which tile's statistics each k-step holds, who has written them, whether an MMA is reading them
    aux = {k: dict(tile=-1, written=set(), reading=False) for k in ("lse", "delta")}
    PUB = {"lse": range(4, 8), "delta": range(8, 12)}      # warps of query quarter 1 / 2 rewrite the lse / delta terms

    def aux_write(kind, w, tile):
        a = aux[kind]
        check(not a["reading"], f"worker {w} rewrites the {kind} terms while an MMA reads them")
        if a["tile"] != tile:
            a["tile"], a["written"] = tile, set()
        a["written"].add(w)

    def tma():
        sim.tma(kv_full)
        Qst[0]["tile"] = 0; sim.tma(q_full[0])
        for i in range(1, nq):
            st = i % NS
            yield ("wait", q_empty[st], i // NS - 1) if i >= NS else ("now",)
            check(Qst[st]["busy"] == 0, f"Q/dO stage {st} reloaded while {Qst[st]['busy']} MMAs read it")
            Qst[st]["tile"] = i; sim.tma(q_full[st])

    def s_mma(i):
        st = i % NS
        def start():
            check(Qst[st]["tile"] == i, f"S^T({i}) reads stage holding {Qst[st]['tile']}")
            check(i == 0 or (Sreg["tile"] == i - 1 and len(Sreg["loaded"]) == NW), f"S^T({i}) overwrites unread scores")
            check(i == 0 or Pbuf["consumed"] >= i - 1, f"S^T({i}) overwrites P^T({i - 1}) before dV read it")
            check(aux["lse"]["tile"] == i and len(aux["lse"]["written"]) == 4, f"S^T({i}) reads lse terms of tile {aux['lse']['tile']}")
            aux["lse"]["reading"] = True
            Qst[st]["busy"] += 1
        def end():
            Sreg["tile"], Sreg["loaded"] = i, set(); Qst[st]["busy"] -= 1; aux["lse"]["reading"] = False
        sim.mma(320, start, end)
        sim.commit(s_full)

    def dp_mma(i):
        st = i % NS
        def start():
            check(Qst[st]["tile"] == i, f"dP^T({i}) reads stage holding {Qst[st]['tile']}")
            check(i == 0 or (DPreg["tile"] == i - 1 and len(DPreg["loaded"]) == NW), f"dP^T({i}) overwrites unread dP^T")
            check(i == 0 or DSbuf["dk_done"] >= i - 1, f"dP^T({i}) overwrites dS^T({i - 1}) before dK read it")
            check(aux["delta"]["tile"] == i and len(aux["delta"]["written"]) == 4, f"dP^T({i}) reads delta terms of tile {aux['delta']['tile']}")
            aux["delta"]["reading"] = True
            Qst[st]["busy"] += 1
        def end():
            DPreg["tile"], DPreg["loaded"] = i, set(); Qst[st]["busy"] -= 1; aux["delta"]["reading"] = False
        sim.mma(320, start, end)
        sim.commit(dp_full)

    def mma():
        yield ("wait", kv_full, 0)
        yield ("wait", q_full[0], 0)
        yield ("wait", aux_init, 0)
        s_mma(0)
        dp_mma(0)
        for i in range(nq):
            st, more = i % NS, i + 1 < nq
            if more:
                yield ("wait", q_full[(i + 1) % NS], (i + 1) // NS)
            yield ("wait", p_ready, i)
            def dv_start(i=i, st=st):
                check(Pbuf["tile"] == i and len(Pbuf["written"]) == NW, f"dV({i}) reads P^T holding {Pbuf['tile']}")
                Qst[st]["busy"] += 1
            def dv_end(i=i, st=st):
                Pbuf["consumed"] = i; Qst[st]["busy"] -= 1
            sim.mma(256, dv_start, dv_end)
            if more:
                s_mma(i + 1)
            yield ("wait", ds_ready, i)
            if i > 0:
                yield ("wait", dq_free, i - 1)
            def dk_start(i=i, st=st):
                check(DSbuf["tile"] == i and len(DSbuf["written"]) == NW, f"dK({i}) reads dS^T holding {DSbuf['tile']}")
                Qst[st]["busy"] += 1
            def dk_end(i=i, st=st):
                DSbuf["dk_done"] = i; Qst[st]["busy"] -= 1
            def dq_start(i=i):
                check(DSbuf["tile"] == i, f"dQ({i}) reads dS holding {DSbuf['tile']}")
                check(DQ["drained"] >= i - 1, f"dQ({i}) overwrites the undrained partial {DQ['tile']}")
            def dq_end(i=i):
                DSbuf["dq_done"] = i; DQ["tile"] = i
            sim.mma(256, dk_start, dk_end)
            if more:
                dp_mma(i + 1)
            sim.mma(384, dq_start, dq_end)
            sim.commit(dq_full)
            sim.commit(q_empty[st])
        sim.commit(acc_full)

    def worker(w):
        for kind in ("lse", "delta"):
            if w in PUB[kind]:
                aux_write(kind, w, 0)
        yield ("sleep", sim.rng.uniform(5, 80))
        aux_init.arrive(32)
        for i in range(nq):
            if bug == "early_lse_write" and w in PUB["lse"] and i + 1 < nq:
                aux_write("lse", w, i + 1)
            yield ("wait", s_full, i)
            if bug != "early_lse_write" and w in PUB["lse"] and i + 1 < nq:
                aux_write("lse", w, i + 1)      # S^T(i) has completed; S^T(i+1) waits for this warp's p_ready(i) arrival
            check(Sreg["tile"] == i, f"worker {w} loads scores of tile {Sreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(200, 1200))
            Sreg["loaded"].add(w)
            if Pbuf["tile"] != i:
                Pbuf["tile"], Pbuf["written"] = i, set()
            Pbuf["written"].add(w)
            p_ready.arrive(32)
            yield ("wait", dp_full, i)
            if w in PUB["delta"] and i + 1 < nq:
                aux_write("delta", w, i + 1)    # dP^T(i) has completed; dP^T(i+1) waits for this warp's ds_ready(i) arrival
            if i > 0 and bug != "no_dq_full_wait":
                yield ("wait", dq_full, i - 1)
            check(DPreg["tile"] == i, f"worker {w} loads dP^T of tile {DPreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(100, 900))
            DPreg["loaded"].add(w)
            check(DSbuf["dq_done"] >= i - 1, f"dS smem overwritten before dQ({i - 1}) finished")
            if DSbuf["tile"] != i:
                DSbuf["tile"], DSbuf["written"] = i, set()
            DSbuf["written"].add(w)
            ds_ready.arrive(32)
        yield ("wait", acc_full, 0)

    def drain(w):
        for i in range(nq):
            yield ("wait", dq_full, i)
            check(DQ["tile"] == i, f"drain loads dQ partial {DQ['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(30, 150))
            DQ["drained"] = max(DQ["drained"], i) if w == 3 else DQ["drained"]  # last drain warp marks the tile (approximation)
            dq_free.arrive(32)
            yield ("sleep", sim.rng.uniform(100, 2000))   # staging + TMA reduce-add

    sim.spawn("tma", tma()); sim.spawn("mma", mma())
    for w in range(NW):
        sim.spawn(f"worker{w}", worker(w))
    for w in range(4):
        sim.spawn(f"drain{w}", drain(w))
    sim.run()


KERNELS = {
    "attn_fwd4_kernel": lambda n, seed, bug=None: run_fwd4(n, seed, bug),
    "attn_bwd4_kernel": lambda n, seed, bug=None: run_bwd4(n, seed, bug),
}


def sweep(name, sizes=range(1, 9), schedules=200, bug=None):
    for n in sizes:
        for seed in range(schedules):
            KERNELS[name](n, seed * 7919 + n, bug)


if __name__ == "__main__":
    for name in KERNELS:
        sweep(name)
        print(f"{name}: 8 sizes x 200 random schedules, no deadlock, no aliasing, no hazard")
