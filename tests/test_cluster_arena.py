"""Host logic of the dense cluster kernel (csrc/dense_cluster.cu): the liveness-based layout of the per-CTA shared-memory arena
(tinydiff.dense.cluster_arena_layout).  The rule under test: while a CTA pushes the output of the Linear that ends epoch e, a slower
peer may still be anywhere in epoch e, so two buffers may share a region only if one of them was last touched in an epoch
strictly before the one the other is first written in.  Pure Python: runs without a GPU."""
import random

import pytest
import torch

from tinydiff.dense import DenseEngine, cluster_arena_layout


def _check(recs, ld, rows, base):
    epoch, e = [], 0
    for r in recs:
        epoch.append(e)
        if r["kind"] == 0 and not r["global"]:
            e += 1
    first, last = {}, {}
    for j, r in enumerate(recs):
        for m in r["reads"] + [r["write"]]:
            last[m[0]] = j
        first.setdefault(r["write"][0], j)
    names = [n for n in base]
    for n in names:
        assert base[n] % 4 == 0 and base[n] >= 0
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            overlap = base[a] < base[b] + rows * ld[b] and base[b] < base[a] + rows * ld[a]
            if not overlap:
                continue
            # disjoint in time, with an epoch boundary in between
            assert epoch[last[a]] < epoch[first[b]] or epoch[last[b]] < epoch[first[a]], (a, b)
    # every record's operands are placed (reads of a buffer nobody wrote would be a builder bug)
    for r in recs:
        for m in r["reads"] + [r["write"]]:
            assert m[0] == "eps" or m[0] in base


def _random_tape(rng, n_ops):
    widths, recs, alive = {"in": rng.choice([20, 64])}, [], ["in"]
    recs.append({"kind": 4, "reads": [], "write": ("in", 0, widths["in"]), "global": False})
    for i in range(n_ops):
        name = f"b{i}"
        widths[name] = rng.choice([20, 64, 128, 256, 512, 1024])
        kind = rng.choice([0, 0, 0, 1, 2])
        src = rng.choice(alive[-4:])
        reads = [(src, 0, widths[src])]
        if kind == 0 and rng.random() < 0.4:
            res = rng.choice(alive)
            reads.append((res, 0, widths[res]))
        recs.append({"kind": kind, "reads": reads, "write": (name, 0, widths[name]), "global": False})
        alive.append(name)
    widths["eps"] = 20
    recs.append({"kind": 0, "reads": [(alive[-1], 0, widths[alive[-1]])], "write": ("eps", 0, 20), "global": True})
    return recs, {n: (w + 3) // 4 * 4 for n, w in widths.items()}


@pytest.mark.parametrize("seed", range(25))
def test_random_tapes_never_alias_within_an_epoch(seed):
    rng = random.Random(seed)
    recs, ld = _random_tape(rng, rng.randint(3, 30))
    for rows in (8, 9):
        base = cluster_arena_layout(recs, ld, rows, 1 << 20)
        assert base is not None
        _check(recs, ld, rows, base)


def test_arena_too_small_is_reported():
    recs, ld = _random_tape(random.Random(1), 10)
    assert cluster_arena_layout(recs, ld, 9, 64) is None


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_model_tapes_fit_the_kernel_arena(name):
    """The two reference models, 9 rows per cluster: layout valid and within the 18432 floats of CK_ARENA_FLOATS, with the reuse
    that makes it fit (the buffers alone are 2-3x the arena)."""
    import importlib
    mod = importlib.import_module(f"tinydiff.{name}")
    kw = {"dropout": 0.0} if name == "diffusion_transformer" else {}
    m = mod.NoiseModel(**kw).eval()
    e = DenseEngine(m, 128, torch.device("cpu"), False, m.in_dim, m.emb_mode)
    m._declare(e)
    recs, written = [], set()
    for op in e._ops:                                   # the builder's abstract records, without fusion / merging
        k = op["kind"]
        if k == "time":
            recs.append({"kind": 3, "reads": [], "write": op["out"], "global": False})
        elif k == "linear":
            recs.append({"kind": 0, "reads": [x for x in (op["x"], op["res"]) if x is not None], "write": op["out"],
                         "global": op["out"][0] == "eps"})
        elif k == "bn":
            recs.append({"kind": 2, "reads": [op["x"]], "write": op["out"], "global": False})
        elif k == "ln":
            recs.append({"kind": 1, "reads": [op["x"]], "write": op["out"], "global": False})
    loads = []
    for r in recs:
        for x in r["reads"]:
            if x[0] not in written and x[0] not in [l["write"][0] for l in loads]:
                loads.append({"kind": 4, "reads": [], "write": e.full(x[0]), "global": False})
        written.add(r["write"][0])
    recs = loads + recs
    ld = {n: (w + 3) // 4 * 4 for n, w in e.widths.items()}
    base = cluster_arena_layout(recs, ld, 9, 18432)
    assert base is not None
    _check(recs, ld, 9, base)
    assert sum(9 * ld[n] for n in base) > 18432, "without reuse the buffers would not fit: the layout must be reusing regions"
