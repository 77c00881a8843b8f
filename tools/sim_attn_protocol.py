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
# ================================================================================================ backward (bwd2 / bwd3)
def run_bwd(nq, seed, transposed, bug=None):
    """attn_bwd2_kernel (transposed=False) / attn_bwd3_kernel (transposed=True): nq query tiles; 8 worker warps, 4 drain warps."""
    sim = Sim(seed)
    B = lambda n, c: sim.barrier(n, c)
    kv_full, s_full, dp_full = B("kv_full", 1), B("s_full", 1), B("dp_full", 1)
    p_ready, ds_ready, dq_full, dq_free, acc_full = B("p_ready", 256), B("ds_ready", 256), B("dq_full", 1), B("dq_free", 128), B("acc_full", 1)
    p_free = B("p_free", 1)
    q_full, q_empty = [B("q_full0", 1), B("q_full1", 1)], [B("q_empty0", 1), B("q_empty1", 1)]
    Sreg, DPreg = dict(tile=-1, loaded=set()), dict(tile=-1, loaded=set())
    Pbuf = dict(tile=-1, written=set(), consumed=-1)    # bwd2: smem P; bwd3: P^T in TMEM over the score columns
    DSbuf = dict(tile=-1, written=set(), dk_done=-1, dq_done=-1)
    DQ = dict(tile=-1, drained=-1)
    Qst = [dict(tile=-1, busy=0) for _ in range(2)]

    def tma():
        sim.tma(kv_full)
        Qst[0]["tile"] = 0; sim.tma(q_full[0])
        for i in range(1, nq):
            st = i & 1
            yield ("wait", q_empty[st], (i >> 1) - 1) if i >= 2 else ("now",)
            check(Qst[st]["busy"] == 0, f"Q/dO stage {st} reloaded while {Qst[st]['busy']} MMAs read it")
            Qst[st]["tile"] = i; sim.tma(q_full[st])

    def s_mma(i):
        st = i & 1
        def start():
            check(Qst[st]["tile"] == i, f"S({i}) reads stage holding {Qst[st]['tile']}")
            check(i == 0 or (Sreg["tile"] == i - 1 and len(Sreg["loaded"]) == 8), f"S({i}) overwrites unread scores")
            if transposed:
                check(i == 0 or Pbuf["consumed"] >= i - 1, f"S^T({i}) overwrites P^T({i - 1}) before dV read it")
            Qst[st]["busy"] += 1
        def end():
            Sreg["tile"], Sreg["loaded"] = i, set(); Qst[st]["busy"] -= 1
        sim.mma(256, start, end)
        sim.commit(s_full)

    def dp_mma(i):
        st = i & 1
        def start():
            check(Qst[st]["tile"] == i, f"dP({i}) reads stage holding {Qst[st]['tile']}")
            check(i == 0 or (DPreg["tile"] == i - 1 and len(DPreg["loaded"]) == 8), f"dP({i}) overwrites unread dP")
            if transposed:
                check(i == 0 or DSbuf["dk_done"] >= i - 1, f"dP^T({i}) overwrites dS^T({i - 1}) before dK read it")
            Qst[st]["busy"] += 1
        def end():
            DPreg["tile"], DPreg["loaded"] = i, set(); Qst[st]["busy"] -= 1
        sim.mma(256, start, end)
        sim.commit(dp_full)

    def mma():
        yield ("wait", kv_full, 0)
        yield ("wait", q_full[0], 0)
        s_mma(0)
        dp_mma(0)
        for i in range(nq):
            st, more = i & 1, i + 1 < nq
            yield ("wait", p_ready, i)
            if more:
                yield ("wait", q_full[st ^ 1], (i + 1) >> 1)
            def dv_start(i=i, st=st):
                check(Pbuf["tile"] == i and len(Pbuf["written"]) == 8, f"dV({i}) reads P holding {Pbuf['tile']}")
                Qst[st]["busy"] += 1
            def dv_end(i=i, st=st):
                Pbuf["consumed"] = i; Qst[st]["busy"] -= 1
            if transposed:
                sim.mma(256, dv_start, dv_end)
                if more:
                    s_mma(i + 1)
            else:
                if more:
                    s_mma(i + 1)
                sim.mma(384, dv_start, dv_end)
                sim.commit(p_free)
            yield ("wait", ds_ready, i)
            if i > 0:
                yield ("wait", dq_free, i - 1)
            def dk_start(i=i, st=st):
                check(DSbuf["tile"] == i and len(DSbuf["written"]) == 8, f"dK({i}) reads dS holding {DSbuf['tile']}")
                Qst[st]["busy"] += 1
            def dk_end(i=i, st=st):
                DSbuf["dk_done"] = i; Qst[st]["busy"] -= 1
            def dq_start(i=i):
                check(DSbuf["tile"] == i, f"dQ({i}) reads dS holding {DSbuf['tile']}")
                check(DQ["drained"] >= i - 1, f"dQ({i}) overwrites the undrained partial {DQ['tile']}")
            def dq_end(i=i):
                DSbuf["dq_done"] = i; DQ["tile"] = i
            if transposed:
                sim.mma(256, dk_start, dk_end)
                if more:
                    dp_mma(i + 1)
                sim.mma(384, dq_start, dq_end)
            else:
                if more:
                    dp_mma(i + 1)
                sim.mma(384, dk_start, dk_end)
                sim.mma(384, dq_start, dq_end)
            sim.commit(dq_full)
            sim.commit(q_empty[st])
        sim.commit(acc_full)

    def worker(w):
        for i in range(nq):
            if transposed:
                yield ("sleep", sim.rng.uniform(10, 80))  # publish the column statistics, named barrier
            yield ("wait", s_full, i)
            check(Sreg["tile"] == i, f"worker {w} loads scores of tile {Sreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(300, 1500))
            Sreg["loaded"].add(w)
            if not transposed:
                if i > 0:
                    yield ("wait", p_free, i - 1)
                check(Pbuf["consumed"] >= i - 1, f"P smem overwritten before dV({i - 1}) finished")
            if Pbuf["tile"] != i:
                Pbuf["tile"], Pbuf["written"] = i, set()
            Pbuf["written"].add(w)
            p_ready.arrive(32)
            yield ("wait", dp_full, i)
            if i > 0 and bug != "no_dq_full_wait":
                yield ("wait", dq_full, i - 1)
            check(DPreg["tile"] == i, f"worker {w} loads dP of tile {DPreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(150, 800))
            DPreg["loaded"].add(w)
            check(DSbuf["dq_done"] >= i - 1, f"dS smem overwritten before dQ({i - 1}) finished")
            if not transposed:
                check(DSbuf["dk_done"] >= i - 1, f"dS smem overwritten before dK({i - 1}) finished")
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
            yield ("sleep", sim.rng.uniform(100, 600))    # staging + TMA reduce-add

    sim.spawn("tma", tma()); sim.spawn("mma", mma())
    for w in range(8):
        sim.spawn(f"worker{w}", worker(w))
    for w in range(4):
        sim.spawn(f"drain{w}", drain(w))
    sim.run()


# ================================================================================================ backward, split X / Y roles (bwd5)
def run_bwd5(nq, seed, bug=None):
    """attn_bwd5_kernel: nq query tiles; 8 X warps (scores -> P^T), 8 Y warps (P^T, dP^T -> dS^T), 4 dQ drain warps, 3-stage
    Q/dO ring, P^T in its own tensor-memory columns, column statistics triple-buffered per role."""
    sim = Sim(seed)
    B = lambda n, c: sim.barrier(n, c)
    kv_full, s_full, s_free, dp_full = B("kv_full", 1), B("s_full", 1), B("s_free", 256), B("dp_full", 1)
    p_ready, p_free, ds_ready = B("p_ready", 256), B("p_free", 257), B("ds_ready", 256)
    dq_full, dq_free, acc_full = B("dq_full", 1), B("dq_free", 128), B("acc_full", 1)
    NS = 3
    q_full, q_empty = [B(f"q_full{k}", 1) for k in range(NS)], [B(f"q_empty{k}", 1) for k in range(NS)]
    Sreg = dict(tile=-1, loaded=set())                      # S^T columns: which tile, which X warps hold it in registers
    Pbuf = dict(tile=-1, written=set(), y_loaded=set(), dv_done=-1)
    DPreg = dict(tile=-1, loaded=set())                     # dP^T columns (dS^T is packed over them by the Y warps)
    DS = dict(tile=-1, written=set(), dk_done=-1, dq_done=-1)
    DQ = dict(tile=-1, drained=-1)
    Qst = [dict(tile=-1, busy=0) for _ in range(NS)]
    stat = {role: [dict(tile=-1, written=set()) for _ in range(3)] for role in "xy"}
    reading = {role: {} for role in "xy"}                   # warp -> tile whose statistics buffer it is reading

    def publish(role, w, tile):
        b = stat[role][tile % 3]
        for ow, ot in reading[role].items():
            check(ot % 3 != tile % 3 or ot == tile, f"{role}{w} overwrites the statistics buffer {role}{ow} reads for tile {ot}")
        if b["tile"] != tile:
            b["tile"], b["written"] = tile, set()
        b["written"].add(w)

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
            check(i == 0 or (Sreg["tile"] == i - 1 and len(Sreg["loaded"]) == 8), f"S^T({i}) overwrites scores not yet in registers")
            Qst[st]["busy"] += 1
        def end():
            Sreg["tile"], Sreg["loaded"] = i, set(); Qst[st]["busy"] -= 1
        sim.mma(256, start, end)
        sim.commit(s_full)

    def dp_mma(i):
        st = i % NS
        def start():
            check(Qst[st]["tile"] == i, f"dP^T({i}) reads stage holding {Qst[st]['tile']}")
            check(i == 0 or (DPreg["tile"] == i - 1 and len(DPreg["loaded"]) == 8), f"dP^T({i}) overwrites unread dP^T")
            check(i == 0 or DS["dk_done"] >= i - 1, f"dP^T({i}) overwrites dS^T({i - 1}) before dK read it")
            Qst[st]["busy"] += 1
        def end():
            DPreg["tile"], DPreg["loaded"] = i, set(); Qst[st]["busy"] -= 1
        sim.mma(256, start, end)
        sim.commit(dp_full)

    def mma():
        yield ("wait", kv_full, 0)
        yield ("wait", q_full[0], 0)
        s_mma(0)
        dp_mma(0)
        for i in range(nq):
            st, more = i % NS, i + 1 < nq
            if more:
                yield ("wait", q_full[(i + 1) % NS], (i + 1) // NS)
                if bug != "no_s_free_wait":
                    yield ("wait", s_free, i)
                s_mma(i + 1)
            yield ("wait", p_ready, i)
            def dv_start(i=i, st=st):
                check(Pbuf["tile"] == i and len(Pbuf["written"]) == 8, f"dV({i}) reads P^T holding {Pbuf['tile']}")
                Qst[st]["busy"] += 1
            def dv_end(i=i, st=st):
                Pbuf["dv_done"] = i; Qst[st]["busy"] -= 1
            sim.mma(256, dv_start, dv_end)
            sim.commit(p_free)
            yield ("wait", ds_ready, i)
            if i > 0:
                yield ("wait", dq_free, i - 1)
            def dk_start(i=i, st=st):
                check(DS["tile"] == i and len(DS["written"]) == 8, f"dK({i}) reads dS^T holding {DS['tile']}")
                Qst[st]["busy"] += 1
            def dk_end(i=i, st=st):
                DS["dk_done"] = i; Qst[st]["busy"] -= 1
            def dq_start(i=i):
                check(DS["tile"] == i, f"dQ({i}) reads dS holding {DS['tile']}")
                check(DQ["drained"] >= i - 1, f"dQ({i}) overwrites the undrained partial {DQ['tile']}")
            def dq_end(i=i):
                DS["dq_done"] = i; DQ["tile"] = i
            sim.mma(256, dk_start, dk_end)
            if more:
                dp_mma(i + 1)
            sim.mma(384, dq_start, dq_end)
            sim.commit(dq_full)
            sim.commit(q_empty[st])
        sim.commit(acc_full)

    def xwarp(w):
        if w < 4:
            publish("x", w, 0)
        yield ("sleep", sim.rng.uniform(5, 60))          # named barrier of the role (modelled as a delay: all publish before any reads)
        for i in range(nq):
            yield ("wait", s_full, i)
            if w < 4 and i + 1 < nq:
                publish("x", w, i + 1)
            reading["x"][w] = i
            check(Sreg["tile"] == i, f"x{w} loads scores of tile {Sreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(80, 300))
            Sreg["loaded"].add(w)
            s_free.arrive(32)
            yield ("sleep", sim.rng.uniform(400, 1800))  # exponentials
            b = stat["x"][i % 3]
            check(b["tile"] == i and len(b["written"]) == 4, f"x{w} read statistics of tile {b['tile']} expecting {i}")
            reading["x"].pop(w)
            if i > 0 and bug != "no_p_free_wait":
                yield ("wait", p_free, i - 1)
            check(i == 0 or (Pbuf["dv_done"] >= i - 1 and (Pbuf["tile"] == i or len(Pbuf["y_loaded"]) == 8)),
                  f"x{w} overwrites P^T({i - 1}) still in use")
            if Pbuf["tile"] != i:
                Pbuf["tile"], Pbuf["written"], Pbuf["y_loaded"] = i, set(), set()
            Pbuf["written"].add(w)
            p_ready.arrive(32)
        yield ("wait", acc_full, 0)

    def ywarp(w):
        if w < 4:
            publish("y", w, 0)
        yield ("sleep", sim.rng.uniform(5, 60))
        for i in range(nq):
            yield ("wait", dp_full, i)
            if w < 4 and i + 1 < nq:
                publish("y", w, i + 1)
            yield ("wait", p_ready, i)
            reading["y"][w] = i
            check(Pbuf["tile"] == i and len(Pbuf["written"]) == 8, f"y{w} loads P^T of tile {Pbuf['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(60, 250))
            Pbuf["y_loaded"].add(w)
            p_free.arrive(32)
            if i > 0 and bug != "no_dq_full_wait":
                yield ("wait", dq_full, i - 1)
            check(DPreg["tile"] == i, f"y{w} loads dP^T of tile {DPreg['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(300, 1400))
            DPreg["loaded"].add(w)
            b = stat["y"][i % 3]
            check(b["tile"] == i and len(b["written"]) == 4, f"y{w} read statistics of tile {b['tile']} expecting {i}")
            reading["y"].pop(w)
            check(DS["dq_done"] >= i - 1, f"dS smem overwritten before dQ({i - 1}) finished")
            if DS["tile"] != i:
                DS["tile"], DS["written"] = i, set()
            DS["written"].add(w)
            ds_ready.arrive(32)
        yield ("wait", acc_full, 0)

    def drain(w):
        for i in range(nq):
            yield ("wait", dq_full, i)
            check(DQ["tile"] == i, f"drain loads dQ partial {DQ['tile']} expecting {i}")
            yield ("sleep", sim.rng.uniform(30, 150))
            DQ["drained"] = max(DQ["drained"], i) if w == 3 else DQ["drained"]  # last drain warp marks the tile (approximation)
            dq_free.arrive(32)
            yield ("sleep", sim.rng.uniform(100, 2500))   # staging + TMA reduce-add

    sim.spawn("tma", tma()); sim.spawn("mma", mma())
    for w in range(8):
        sim.spawn(f"x{w}", xwarp(w)); sim.spawn(f"y{w}", ywarp(w))
    for w in range(4):
        sim.spawn(f"drain{w}", drain(w))
    sim.run()


KERNELS = {
    "attn_fwd4_kernel": lambda n, seed, bug=None: run_fwd4(n, seed, bug),
    "attn_bwd3_kernel": lambda n, seed, bug=None: run_bwd(n, seed, True, bug),
    "attn_bwd5_kernel": lambda n, seed, bug=None: run_bwd5(n, seed, bug),
}


def sweep(name, sizes=range(1, 9), schedules=200, bug=None):
    for n in sizes:
        for seed in range(schedules):
            KERNELS[name](n, seed * 7919 + n, bug)


if __name__ == "__main__":
    for name in KERNELS:
        sweep(name)
        print(f"{name}: 8 sizes x 200 random schedules, no deadlock, no aliasing, no hazard")
