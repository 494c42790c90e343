"""The chain kernel queues |dH| per macro step and works the queue off once per transition,
one update per lane (walnuts_b200/csrc/chain_kernel.cuh: adam_flush).  This restates both
schedules in plain IEEE doubles and checks the claim the kernel relies on: splitting each
Adam update (adam.hpp:70-86) into the sequential recurrences (t, beta powers, m, v, x) and
the independent per-update part (exp, divisions, square root, decayed rate) gives the same
bits as updating one at a time."""
import math

import numpy as np

LR, B1, B2, EPS, DECAY, TARGET = 0.05, 0.8, 0.9, 1e-4, 0.5, 0.8


def sequential(state, dH):
    t, b1p, b2p, m, v, x = state
    for d in dH:
        alpha = math.exp(-d)
        t += 1.0
        b1p *= B1
        b2p *= B2
        grad = TARGET - alpha
        m = B1 * m + (1 - B1) * grad
        v = B2 * v + (1 - B2) * grad * grad
        m_hat = m / (1 - b1p)
        v_hat = v / (1 - b2p)
        decayed = LR / math.pow(t, DECAY)
        denom = math.sqrt(v_hat) + EPS
        x -= decayed * m_hat / denom
    return t, b1p, b2p, m, v, x


def queued(state, dH, lanes=32):
    t, b1p, b2p, m, v, x = state
    for base in range(0, len(dH), lanes):
        chunk = dH[base:base + lanes]
        alpha = [math.exp(-d) for d in chunk]              # one per lane
        mine = []
        for a in alpha:                                    # every lane runs the recurrence
            t += 1.0
            b1p *= B1
            b2p *= B2
            grad = TARGET - a
            m = B1 * m + (1 - B1) * grad
            v = B2 * v + (1 - B2) * grad * grad
            mine.append((t, b1p, b2p, m, v))               # lane j keeps step j
        terms = []
        for (tj, b1j, b2j, mj, vj) in mine:                # one per lane
            m_hat = mj / (1 - b1j)
            v_hat = vj / (1 - b2j)
            decayed = LR / math.pow(tj, DECAY)
            denom = math.sqrt(v_hat) + EPS
            terms.append(decayed * m_hat / denom)
        for term in terms:                                 # in order, as the shuffles do
            x -= term
    return t, b1p, b2p, m, v, x


def test_queued_adam_equals_one_at_a_time_bitwise():
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 32, 33, 64, 100, 1023):
        dH = list(np.abs(rng.standard_cauchy(n)) * rng.choice([1e-3, 0.1, 5.0], n))
        state = (float(rng.integers(0, 50)), 0.8 ** 3, 0.9 ** 3, rng.normal(), abs(rng.normal()),
                 math.log(0.3))
        assert sequential(state, dH) == queued(state, dH)
