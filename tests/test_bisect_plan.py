"""Host logic of the bisection's speculation plan (cpb_bisect_plan; no GPU needed).

The reference's loop (BisectCostBottleneckSplitter.jl:41-60) visits one root-to-leaf path of the bisection tree.  A round
probes a set of tree nodes concurrently and then walks them by feasibility; whatever the set, replaying the rounds must
reproduce the sequential loop's final bracket exactly -- the set only changes how many rounds are needed."""
import math

import numpy as np

import chainb200 as cp


def sequential(c_lo, c_hi, eps, cstar):
    probes = 0
    while c_lo * (1 + eps) < c_hi:
        c = (c_lo + c_hi) / 2
        probes += 1
        if c >= cstar:
            c_hi = c
        else:
            c_lo = c
    return c_lo, c_hi, probes


def by_rounds(c_lo, c_hi, eps, cstar, nodes, ub):
    c_lo0, c_hi0 = c_lo, c_hi
    rounds = probes = 0
    while c_lo * (1 + eps) < c_hi:
        ids = cp.bisect_plan(c_lo, c_hi, eps, nodes, c_lo0, c_hi0, ub)
        used = [i for i in ids if i >= 0]
        assert used and used[0] == 0 and len(set(used)) == len(used)          # the root is always probed
        assert all(i == 0 or (i - 1) // 2 in used for i in used)               # closed under taking parents
        heap = 0
        while heap in used and c_lo * (1 + eps) < c_hi:                        # the walk of k_bisect_advance
            c = (c_lo + c_hi) / 2
            probes += 1
            if c >= cstar:
                c_hi, heap = c, 2 * heap + 1
            else:
                c_lo, heap = c, 2 * heap + 2
        rounds += 1
        assert rounds < 200
    return c_lo, c_hi, probes, rounds


def test_flat_prior_is_the_complete_tree_in_heap_order():
    # no upper bound and a bracket the log-uniform prior cannot be built on (c_lo = 0): masses are widths -> BFS order
    assert cp.bisect_plan(0.0, 1024.0, 1e-6, 15, 0.0, 1024.0, 0.0) == list(range(15))
    assert cp.bisect_plan(0.0, 1024.0, 1e-6, 7, 0.0, 1024.0, 0.0) == list(range(7))


def test_plan_never_changes_the_bracket():
    rng = np.random.default_rng(21)
    for _ in range(300):
        K = int(rng.choice([2, 8, 64, 1024]))
        eps = float(rng.choice([0.1, 0.01, 0.001]))
        c_lo = float(rng.integers(1, 10 ** 7))
        c_hi = c_lo * K * float(rng.uniform(0.9, 1.1))
        cstar = c_lo * float(np.exp(rng.uniform(0, math.log(K))))
        ub = float(rng.choice([0.0, cstar * 1.0001, cstar * 1.02, cstar * 1.3, cstar * 3, c_hi]))
        want = sequential(c_lo, c_hi, eps, cstar)
        for nodes in (1, 3, 15, 30):
            got = by_rounds(c_lo, c_hi, eps, cstar, nodes, min(ub, c_hi))
            assert got[:3] == want, (K, eps, nodes, ub / cstar)


def test_tight_upper_bound_saves_rounds():
    # K = 64, eps = 0.01, optimum 8.6 x c_lo (config 2's regime): 10 probes
    c_lo, c_hi, eps, cstar = 1.0e6, 64.0e6, 0.01, 8.6e6
    probes = sequential(c_lo, c_hi, eps, cstar)[2]
    full_tree = math.ceil(probes / 4)                                   # complete 4-level trees
    no_bound = by_rounds(c_lo, c_hi, eps, cstar, 15, 0.0)[3]
    tight = by_rounds(c_lo, c_hi, eps, cstar, 15, cstar * 1.0001)[3]
    assert tight == 1 and tight < no_bound <= full_tree
    # a useless bound (3 x the optimum) must not cost more than one round over the prior without a bound
    assert by_rounds(c_lo, c_hi, eps, cstar, 15, cstar * 3)[3] <= no_bound + 1


def test_prewalk_keeps_the_reference_result():
    """cpb_bisect_prewalk walks the loop's leading probes that an upper bound ub >= c* settles (all feasible).  Whatever the
    bracket -- including a c_lo ABOVE the optimum, which bound_stripe produces for the envelope model (found by the fuzzer) --
    the loop continued from the prewalked state must end in the same bracket as the sequential loop, and the loop's LAST
    feasible probe (whose split vector is the result) must not be among the prewalked ones."""
    rng = np.random.default_rng(77)
    for _ in range(4000):
        eps = float(rng.choice([0.3, 0.1, 0.01, 0.001]))
        c_lo = float(rng.uniform(1, 1e7))
        c_hi = c_lo * float(np.exp(rng.uniform(0, math.log(2000))))
        # the optimum anywhere around the bracket: below c_lo (invalid lower bound), inside, or at c_hi
        cstar = c_lo * float(np.exp(rng.uniform(-0.5, math.log(c_hi / c_lo))))
        ub = cstar * float(rng.choice([1.0, 1.0001, 1.01, 1.3, 3.0]))
        # sequential loop, recording the probes
        lo, hi, seq = c_lo, c_hi, []
        while lo * (1 + eps) < hi:
            c = (lo + hi) / 2
            ok = c >= cstar
            seq.append((c, ok))
            if ok:
                hi = c
            else:
                lo = c
        hi2, walked = cp.bisect_prewalk(c_lo, c_hi, eps, ub)
        assert 0 <= walked <= len(seq)
        assert all(ok for _, ok in seq[:walked])                      # only feasible probes are settled
        assert hi2 == (seq[walked - 1][0] if walked else c_hi)          # ... with the loop's own thresholds
        feas = [i for i, (_, ok) in enumerate(seq) if ok]
        if feas:
            assert feas[-1] >= walked, (c_lo, c_hi, eps, cstar, ub)    # the last feasible probe still runs for real
