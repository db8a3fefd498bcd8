"""The specialised state kernel replaces idx / T by umulhi(idx, M), M = floor(2^32 / T) + 1, whenever the host can prove the
multiply exact for the launch (engine.cu: finish_state_args; StateKernelArgs::t_magic).  The proof: M T = 2^32 + r with
0 < r <= T, so idx M / 2^32 = idx / T + idx r / (T 2^32), and the floor is that of idx / T as long as idx r < 2^32 — which
num_gen T^2 < 2^32 guarantees for every idx < num_gen T.  Checked here on the integers, for every trajectory length the
engine takes and at the edge of the condition."""
import numpy as np


def _magic(T, num_gen):
    """finish_state_args' rule, restated."""
    return (1 << 32) // T + 1 if T >= 2 and num_gen * T * T < (1 << 32) else 0


def test_multiply_equals_division_wherever_the_host_enables_it():
    rng = np.random.default_rng(0)
    for T in range(2, 257):
        # the largest rollout count the rule accepts for this T, and the shipped ones
        n_max = ((1 << 32) - 1) // (T * T)
        for num_gen in {1, 64, 4096, n_max}:
            M = _magic(T, num_gen)
            if num_gen > n_max:
                assert M == 0
                continue
            assert M != 0 and M < (1 << 32)
            top = num_gen * T
            idx = np.unique(np.concatenate([
                np.arange(0, min(top, 4 * T)), np.arange(max(0, top - 4 * T), top),
                rng.integers(0, top, size=2000), (np.arange(1, min(num_gen, 3000)) * T) - 1, np.arange(0, min(num_gen, 3000)) * T,
            ])).astype(np.uint64)
            idx = idx[idx < top]
            q = (idx * np.uint64(M)) >> np.uint64(32)
            np.testing.assert_array_equal(q, idx // np.uint64(T))


def test_rule_refuses_what_it_cannot_prove():
    assert _magic(1, 10) == 0                       # T = 1: the multiplier would not fit 32 bits
    assert _magic(256, 65536) == 0                  # 2^16 rollouts of 2^8 steps: num_gen T^2 = 2^32
    assert _magic(256, 65535) != 0
    # just beyond the bound the multiply can really be wrong: the refusal is not only caution
    T = 255
    M = (1 << 32) // T + 1
    r = M * T - (1 << 32)
    bad = [i for i in range((1 << 32) // r, (1 << 32) // r + 4 * T) if ((i * M) >> 32) != i // T]
    assert bad, "expected a counterexample just past idx r = 2^32"
