"""The oracle's verifier chain over a whole proof in the flat layouts of include/swirl_b200.h
(verifier/mod.rs order: transcript prefix, batch constraints, stacked reduction, WHIR).
Test infrastructure: works at any trace size because the verifier never reads the traces (only
their shapes); used by tests/test_prove.py and tools/run_configs.py."""
import numpy as np

import airs as A


def mont1(x):
    return np.array([A.to_mont(x)], dtype=np.uint32)


def verify(oracle, l_skip, n_stack, log_blowup, D, logup_pow, whir, vk_pre_hash, airs_by_id, is_required, root,
           pre_cached_roots, bc, stacking, whir_proof):
    """airs_by_id: airs.Air list (traces may be zero-filled: only shapes matter); pre_cached_roots[air_id] =
    (preprocessed root | None, [cached roots]).  Returns (ok, stage that failed | None)."""
    order = sorted(range(len(airs_by_id)), key=lambda i: (-airs_by_id[i].height, i))
    sa = [airs_by_id[i] for i in order]
    st = np.zeros(18, np.uint32)
    oracle.sponge_observe(st, vk_pre_hash)
    oracle.sponge_observe(st, root)
    for air_id, a in enumerate(airs_by_id):
        if not is_required[air_id]:
            oracle.sponge_observe(st, mont1(1))
        prep_root, cached_roots = pre_cached_roots[air_id]
        oracle.sponge_observe(st, prep_root if prep_root is not None else mont1(a.height.bit_length() - 1))
        for c in cached_roots:
            oracle.sponge_observe(st, c)
        oracle.sponge_observe(st, a.public_values)
    n_max = max(sa[0].height.bit_length() - 1 - l_skip, 0)
    ok, r = oracle.bc_verify(st, l_skip, D, logup_pow, A.flatten(sa), len(sa), n_max, bc)
    if not ok:
        return False, "batch_constraints"
    n_open = sum((a.common_main[2] + sum(m[2] for m in a.cached) + (a.preprocessed[2] if a.preprocessed is not None else 0))
                 * (2 if a.need_rot else 1) for a in sa)
    op = bc[-4 * n_open:].reshape(-1, 4)
    pos, per_air = 0, []
    for a in sa:
        parts = []
        for m in [a.common_main] + ([a.preprocessed] if a.preprocessed is not None else []) + a.cached:
            n = m[2] * (2 if a.need_rot else 1)
            parts.append(op[pos:pos + n])
            pos += n
        per_air.append(parts)
    zero = np.zeros(4, np.uint32)

    def pairs(part, rot):
        if rot:
            return [np.concatenate([part[2 * i], part[2 * i + 1]]) for i in range(len(part) // 2)]
        return [np.concatenate([c, zero]) for c in part]

    t_claims = [p_ for a, parts in zip(sa, per_air) for p_ in pairs(parts[0], a.need_rot)]
    for a, parts in zip(sa, per_air):
        for part in parts[1:]:
            t_claims += pairs(part, a.need_rot)
    shape = lambda m, rot: (np.zeros(0, np.uint32), m[1], m[2], rot)
    commits = [[shape(a.common_main, a.need_rot) for a in sa]]
    roots = [root]
    for i, a in zip(order, sa):
        prep_root, cached_roots = pre_cached_roots[i]
        for m, rt in zip(([a.preprocessed] if a.preprocessed is not None else []) + a.cached,
                         ([prep_root] if prep_root is not None else []) + list(cached_roots)):
            commits.append([shape(m, a.need_rot)])
            roots.append(rt)
    ok, u = oracle.stacked_reduction_verify(st, l_skip, n_stack, commits, np.array(t_claims), r, stacking)
    if not ok:
        return False, "stacked_reduction"
    u_cube = [u[0]]
    for _ in range(l_skip - 1):
        u_cube.append(oracle.ef_mul(u_cube[-1], u_cube[-1]))
    u_cube = np.array(u_cube + list(u[1:]), dtype=np.uint32)
    n0 = (2 * ((1 << l_skip) - 1) + 1) * 4 + n_stack * 8
    openings = stacking[n0:].reshape(-1, 4)
    H = 1 << (l_skip + n_stack)
    widths = [(sum(max(h, 1 << l_skip) * w for _, h, w, _ in c) + H - 1) // H for c in commits]
    if not oracle.whir_verify(st, l_skip, n_stack, log_blowup, whir, whir_proof, widths, openings, np.array(roots), u_cube):
        return False, "whir"
    return True, st
