"""Model check of the peer-memory window protocol (nw_halo.inc: p2p_next,
p2p_pull_begin / p2p_pull_end, p2p_signal_then_wait).  No GPU.

What the product does per rank and exchange e (same sequence on every rank):

  compute stream   K_e   producing kernel.  Fused push: it stores into slot
                         e mod NSLOT of every neighbour's window at arbitrary
                         moments of its run.
                   P_e   (plain push only) push kernel: stores, then publishes
                         epoch e in the neighbours' flag words.
  communication    L_e   pull kernel: (fused push) publishes epoch e first;
  stream                 waits until every neighbour's flag >= e; reads the own
                         slot e mod NSLOT; completion event.

  orderings the host code issues
    L_e after K_e / P_e            (p2p_pull_begin: event on the compute stream)
    L_e after L_{e-1}              (one communication stream)
    fused K_e after L_{e-2}        (p2p_next: pullRing[(e-2) mod 3])
    plain P_e after L_{e-1}        (p2p_next: lastPull)

The model runs R ranks with random kernel durations (incl. long stalls of one
rank) and a random fused / plain choice per rank and exchange, as a discrete
event simulation with the orderings above and nothing else, and checks

  (I1) no store of epoch e lands in a slot before its owner has finished the
       pull that read the slot's previous content (epoch e - NSLOT);
  (I2) a pull reads its slot only after every neighbour's stores of that epoch
       have completed.

It also shows that the test has teeth: with two slots, or with the fused
kernel waiting for nothing, (I1) is violated in some schedule."""
import random

import pytest


def simulate(n_ranks, n_exch, n_slot, fused_wait, seed):
    """returns the list of (I1)/(I2) violations of one random schedule.
    fused_wait: how many exchanges back the pull is that a fused kernel waits
    for (product: 2); None: it waits for no pull at all."""
    rng = random.Random(seed)
    R, E = n_ranks, n_exch
    fused = [[rng.random() < 0.6 for _ in range(E + 1)] for _ in range(R)]
    # durations: mostly short, sometimes one rank stalls for a long time
    def dur(lo, hi):
        d = rng.uniform(lo, hi)
        if rng.random() < 0.05:
            d += rng.uniform(20, 80)
        return d
    kdur = [[dur(1, 6) for _ in range(E + 1)] for _ in range(R)]
    pdur = [[dur(0.2, 1) for _ in range(E + 1)] for _ in range(R)]
    ldur = [[dur(0.2, 3) for _ in range(E + 1)] for _ in range(R)]
    gap = [[rng.uniform(0, 2) for _ in range(E + 1)] for _ in range(R)]

    INF = float("inf")
    k_start = [[None] * (E + 1) for _ in range(R)]
    k_end = [[None] * (E + 1) for _ in range(R)]
    p_end = [[None] * (E + 1) for _ in range(R)]      # end of push (stores done)
    flag = [[None] * (E + 1) for _ in range(R)]       # time epoch e is published
    l_start = [[None] * (E + 1) for _ in range(R)]
    l_read = [[None] * (E + 1) for _ in range(R)]     # pull starts reading
    l_end = [[0.0] * (E + 1) for _ in range(R)]
    store = [[None] * (E + 1) for _ in range(R)]      # (first, last) store time
    main_free = [0.0] * R
    comm_free = [0.0] * R

    # fixed point over exchanges: L_e of a rank needs the neighbours' flags of
    # e, which need their K_e / L_e start -- iterate exchange by exchange;
    # inside one exchange compute-stream work first, then flags, then pulls
    for e in range(1, E + 1):
        for r in range(R):
            t = main_free[r] + gap[r][e]
            if fused[r][e]:
                if fused_wait is not None and e - fused_wait >= 1:
                    t = max(t, l_end[r][e - fused_wait])
            k_start[r][e] = t
            k_end[r][e] = t + kdur[r][e]
            if fused[r][e]:
                # stores anywhere inside the kernel
                a = rng.uniform(k_start[r][e], k_end[r][e])
                b = rng.uniform(a, k_end[r][e])
                store[r][e] = (a, b)
                p_end[r][e] = k_end[r][e]
                main_free[r] = k_end[r][e]
            else:
                ps = max(k_end[r][e], l_end[r][e - 1])  # plain push after L_{e-1}
                store[r][e] = (ps, ps + pdur[r][e])
                p_end[r][e] = ps + pdur[r][e]
                flag[r][e] = p_end[r][e]                 # push kernel publishes
                main_free[r] = p_end[r][e]
        for r in range(R):
            l_start[r][e] = max(comm_free[r], p_end[r][e])
            if fused[r][e]:
                flag[r][e] = l_start[r][e]               # pull kernel publishes
        for r in range(R):
            l_read[r][e] = max([l_start[r][e]] +
                               [flag[q][e] for q in range(R) if q != r])
            l_end[r][e] = l_read[r][e] + ldur[r][e]
            comm_free[r] = l_end[r][e]

    bad = []
    for e in range(1, E + 1):
        for p in range(R):
            for q in range(R):
                if p == q:
                    continue
                first, last = store[p][e]
                if e - n_slot >= 1 and first < l_end[q][e - n_slot]:
                    bad.append(("I1", e, p, q, first, l_end[q][e - n_slot]))
                if l_read[q][e] < last:
                    bad.append(("I2", e, p, q, last, l_read[q][e]))
    return bad


@pytest.mark.parametrize("n_ranks", [2, 3, 5])
def test_three_slots_and_the_products_orderings_are_safe(n_ranks):
    for seed in range(300):
        bad = simulate(n_ranks, 40, 3, 2, seed)
        assert not bad, (seed, bad[:3])


def test_two_slots_are_not_enough_for_the_relaxed_ordering():
    # fused kernel waits for the pull two exchanges back, windows alternate
    # between two halves: a neighbour's stores of e + 2 can overtake my pull of e
    assert any(simulate(3, 40, 2, 2, seed) for seed in range(300))


def test_a_fused_kernel_must_wait_for_the_pull_two_exchanges_back():
    assert any(simulate(3, 40, 3, None, seed) for seed in range(300))


def test_two_slots_with_the_strict_ordering_are_safe():
    # the protocol before the asynchronous fused exchange: every producer
    # waits for the previous pull (one stream, or lastPull)
    for seed in range(200):
        assert not simulate(3, 40, 2, 1, seed)
