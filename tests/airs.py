"""Test AIRs as symbolic DAGs in the boundary encoding (include/swirl_b200.h: swirl_dag_node), with
their traces.  Mirrors of the reference fixtures:
  FibonacciAir          crates/stark-backend/src/test_utils/dummy_airs/fib_air/{air.rs,trace.rs}
  BenchmarkAir          benchmarks/synthetic/src/bin/uniform_runner.rs:78-113 (assert_bool per column +
                        self-cancelling send/receive pairs on bus 0)
  sender / receiver     the interaction fixtures of crates/backend-tests/src/lib.rs:571-600 (a sender
                        and a receiver of different heights whose LogUp sums cancel)
The DAG node order / dedup of the reference's SymbolicDagBuilder is not reproduced (the DAG is an
input at the boundary, keygen is out of scope); any topologically ordered DAG is valid input."""
import numpy as np

P = 0x78000001
VAR_PREP, VAR_MAIN, VAR_PUBLIC, IS_FIRST, IS_LAST, IS_TRANSITION, CONST, ADD, SUB, NEG, MUL = range(11)


def to_mont(x):
    return int((int(x) % P) * (1 << 32) % P)


class Dag:
    def __init__(self):
        self.nodes, self.cache = [], {}

    def _n(self, *t):
        t = tuple(int(v) for v in t) + (0,) * (4 - len(t))
        if t not in self.cache:
            self.cache[t] = len(self.nodes)
            self.nodes.append(t)
        return self.cache[t]

    def main(self, col, offset=0, part=0): return self._n(VAR_MAIN, col, offset, part)
    def prep(self, col, offset=0): return self._n(VAR_PREP, col, offset)
    def public(self, i): return self._n(VAR_PUBLIC, i)
    def const(self, v): return self._n(CONST, to_mont(v))
    def is_first(self): return self._n(IS_FIRST)
    def is_last(self): return self._n(IS_LAST)
    def is_transition(self): return self._n(IS_TRANSITION)
    def add(self, a, b): return self._n(ADD, a, b)
    def sub(self, a, b): return self._n(SUB, a, b)
    def neg(self, a): return self._n(NEG, a)
    def mul(self, a, b): return self._n(MUL, a, b)


class Air:
    """One present AIR with its trace: the inputs of swirl_air_ctx."""

    def __init__(self, dag, constraints, interactions, constraint_degree, need_rot, common_main, public_values=(),
                 cached=(), preprocessed=None):
        self.nodes = np.array(dag.nodes, dtype=np.uint32).reshape(-1, 4)
        self.constraint_idx = np.array(sorted(set(constraints)), dtype=np.uint32)
        self.interactions = list(interactions)  # (count_node, bus_index, [msg nodes])
        self.constraint_degree, self.need_rot = int(constraint_degree), bool(need_rot)
        self.public_values = np.array([to_mont(v) for v in public_values], dtype=np.uint32)
        self.common_main = common_main  # (mont words col-major flat, height, width)
        self.cached, self.preprocessed = list(cached), preprocessed

    @property
    def height(self):
        return self.common_main[1]

    def mats(self):
        return [self.common_main] + self.cached + ([self.preprocessed] if self.preprocessed is not None else [])


def mont_matrix(cols):
    """list of canonical-int columns -> (Montgomery words, col-major flat, height, width)"""
    h = len(cols[0])
    flat = np.array([to_mont(v) for c in cols for v in c], dtype=np.uint32)
    return (flat, h, len(cols))


def fibonacci(log_n, a=0, b=1):
    n = 1 << log_n
    left, right = [a], [b]
    for _ in range(n - 1):
        left.append(right[-1])
        right.append((left[-2] + right[-1]) % P)
    d = Dag()
    l0, r0, l1, r1 = d.main(0), d.main(1), d.main(0, 1), d.main(1, 1)
    cons = [
        d.mul(d.is_first(), d.sub(l0, d.public(0))),
        d.mul(d.is_first(), d.sub(r0, d.public(1))),
        d.mul(d.is_transition(), d.sub(r0, l1)),
        d.mul(d.is_transition(), d.sub(d.add(l0, r0), r1)),
        d.mul(d.is_last(), d.sub(r0, d.public(2))),
    ]
    return Air(d, cons, [], 2, True, mont_matrix([left, right]), public_values=[a, b, right[-1]])


def benchmark(log_n, cols, constraints, pairs, rng, zero_trace=False):
    n = 1 << log_n
    data = [[0] * n for _ in range(cols)] if zero_trace else [list(rng.integers(0, 2, size=n)) for _ in range(cols)]
    d = Dag()
    one = d.const(1)
    cons = []
    for i in range(constraints):
        x = d.main(i % cols)
        cons.append(d.mul(x, d.sub(x, one)))
    inter = []
    neg_one = d.neg(one)
    for i in range(pairs):
        x = d.main(i % cols)
        inter.append((one, 0, [x]))
        inter.append((neg_one, 0, [x]))
    return Air(d, cons, inter, 2, False, mont_matrix(data))


def sender_receiver(log_send, log_recv, rng, bus=3, balanced=True):
    """sender: columns (mult, v0, v1) sends (v0, v1) with multiplicity mult; receiver: columns (count, v0, v1)
    receives with multiplicity count.  Receiver rows are the distinct messages; counts make the sums cancel."""
    ns, nr = 1 << log_send, 1 << log_recv
    msgs = [(int(rng.integers(0, P)), int(rng.integers(0, P))) for _ in range(nr)]
    pick = rng.integers(0, nr, size=ns)
    mult = rng.integers(0, 3, size=ns)
    counts = [0] * nr
    for p_, m_ in zip(pick, mult):
        counts[p_] += int(m_)
    if not balanced:
        counts[0] += 1
    send_cols = [list(mult), [msgs[p_][0] for p_ in pick], [msgs[p_][1] for p_ in pick]]
    recv_cols = [counts, [m[0] for m in msgs], [m[1] for m in msgs]]
    ds = Dag()
    s_air = Air(ds, [], [(ds.main(0), bus, [ds.main(1), ds.main(2)])], 1, False, mont_matrix(send_cols))
    dr = Dag()
    # a (vacuous but non-trivial) degree-2 constraint keeps the zerocheck path busy too
    c = dr.mul(dr.sub(dr.main(1), dr.main(1)), dr.main(2))
    r_air = Air(dr, [c], [(dr.neg(dr.main(0)), bus, [dr.main(1), dr.main(2)])], 2, False, mont_matrix(recv_cols))
    return s_air, r_air


def dummy_interaction(row_major, is_send, bus=0):
    """DummyInteractionAir::new(1, is_send, bus) (test_utils/dummy_airs/interaction/dummy_interaction_air.rs:94-122) on a
    literal row-major [count, field] table as the reference's interaction tests write them (backend-tests/src/lib.rs:844-1018):
    a send pushes (bus, [field], count), a receive pushes (bus, [field], -count); no constraints."""
    rows = [row_major[i:i + 2] for i in range(0, len(row_major), 2)]
    d = Dag()
    count = d.main(0) if is_send else d.neg(d.main(0))
    return Air(d, [], [(count, bus, [d.main(1)])], 1, False, mont_matrix([[r[0] for r in rows], [r[1] for r in rows]]))


def dummy_interaction_chip(counts, fields, is_send, bus=0, partition=False):
    """DummyInteractionChip (dummy_interaction_air.rs:126-260): rows (count, fields...) padded with zero rows to a power of
    two; with `partition` the count column is the common main and the fields are one cached main (new_with_partition)."""
    n = 1
    while n < len(counts):
        n *= 2
    w = len(fields[0])
    counts = list(counts) + [0] * (n - len(counts))
    fields = [list(f) for f in fields] + [[0] * w] * (n - len(fields))
    d = Dag()
    if partition:
        cnt, msg = d.main(0, 0, 1), [d.main(i, 0, 0) for i in range(w)]  # part 0 = the cached main, last part = common main
        mats = dict(common_main=mont_matrix([counts]), cached=[mont_matrix([[f[i] for f in fields] for i in range(w)])])
    else:
        cnt, msg = d.main(0), [d.main(1 + i) for i in range(w)]
        mats = dict(common_main=mont_matrix([counts] + [[f[i] for f in fields] for i in range(w)]), cached=[])
    return Air(d, [], [(cnt if is_send else d.neg(cnt), bus, msg)], 1, False, mats["common_main"], cached=mats["cached"])


def self_interaction(width, log_height, bus):
    """SelfInteractionAir + SelfInteractionChip (test_utils/dummy_airs/interaction/self_interaction_air.rs:26-86): eight
    interactions on one bus whose messages are the whole local / next row (forward and reversed) with constant, row-sum and
    first-column multiplicities; trace[row][i] = (row + i) mod width, so every pair cancels over the cyclic trace.  As in
    the reference, `next_sum` is (deliberately or not) the sum of the LOCAL row."""
    n = 1 << log_height
    cols = [[(r + i) % width for r in range(n)] for i in range(width)]
    d = Dag()
    local = [d.main(i, 0) for i in range(width)]
    nxt = [d.main(i, 1) for i in range(width)]
    zero, one = d.const(0), d.const(1)
    local_sum = zero
    for v in local:
        local_sum = d.add(local_sum, v)
    next_sum = local_sum
    inter = [
        (one, bus, list(local)), (d.neg(one), bus, list(nxt)),
        (local_sum, bus, list(local)), (d.neg(next_sum), bus, list(nxt)),
        (local[0], bus, list(local)), (d.neg(nxt[0]), bus, list(nxt)),
        (local_sum, bus, list(reversed(local))), (d.neg(next_sum), bus, list(reversed(nxt))),
    ]
    return Air(d, [], inter, 1, True, mont_matrix(cols))


def with_parts(log_n, rng):
    """An AIR with a preprocessed trace and one cached main next to the common main, rotations used:
    prep col p, cached col c, common cols (x, y):  y' = y + p * c  on transitions;  x * (x - 1) = 0;
    is_first * (y - c) = 0."""
    n = 1 << log_n
    p = [int(v) for v in rng.integers(0, P, size=n)]
    c = [int(v) for v in rng.integers(0, P, size=n)]
    x = [int(v) for v in rng.integers(0, 2, size=n)]
    y = [c[0]]
    for i in range(n - 1):
        y.append((y[-1] + p[i] * c[i]) % P)
    d = Dag()
    pv, cv, xv, yv, yn = d.prep(0), d.main(0, 0, 0), d.main(0, 0, 1), d.main(1, 0, 1), d.main(1, 1, 1)
    cons = [
        d.mul(d.is_transition(), d.sub(yn, d.add(yv, d.mul(pv, cv)))),
        d.mul(xv, d.sub(xv, d.const(1))),
        d.mul(d.is_first(), d.sub(yv, cv)),
    ]
    return Air(d, cons, [], 3, True, mont_matrix([x, y]), cached=[mont_matrix([c])], preprocessed=mont_matrix([p]))


def flatten(airs):
    """Arrays for the oracle C API (oracle/capi.cpp: orc_bc_*)."""
    meta, nodes, cidx, inter, msg, pubs, mats = [], [], [], [], [], [], []
    for a in airs:
        meta.append([len(a.nodes), len(a.constraint_idx), len(a.interactions), a.constraint_degree, int(a.need_rot),
                     len(a.public_values), len(a.cached), int(a.preprocessed is not None)])
        nodes.append(a.nodes.reshape(-1))
        cidx.append(a.constraint_idx)
        off = 0
        for cnt, bus, m in a.interactions:
            inter.append([cnt, bus, off, len(m)])
            msg.extend(m)
            off += len(m)
        pubs.append(a.public_values)
        mats.extend(a.mats())
    cat = lambda xs, dt: np.concatenate([np.asarray(x, dtype=dt).reshape(-1) for x in xs]) if xs else np.zeros(0, dt)
    return dict(meta=np.array(meta, dtype=np.uint64).reshape(-1), nodes=cat(nodes, np.uint32), cidx=cat(cidx, np.uint32),
                inter=np.array(inter, dtype=np.uint32).reshape(-1), msg=np.array(msg, dtype=np.uint32), pubs=cat(pubs, np.uint32),
                mats=mats)
