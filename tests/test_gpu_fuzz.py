"""Seeded randomised shapes through every kernel family against the CPU oracle: ragged sizes around the tile /
slab / slice boundaries, k from 1 to beyond N, empty and out-of-vocabulary queries, masked documents."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import fuse as ofuse
from oracle import maxsim as omaxsim
from tests.parity import check_topk_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from legal_rag_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return engine


def _unit(rng, shape):
    x = rng.standard_normal(shape).astype(np.float32)
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _bf16(x):
    t = torch.from_numpy(x).cuda().to(torch.bfloat16)
    return t, t.float().cpu().numpy()


def test_fuzz_dense(eng):
    rng = np.random.default_rng(1001)
    for trial in range(24):
        N = int(rng.choice([1, 7, 255, 256, 257, 1000, 4097, 20011]))
        d = int(rng.choice([8, 64, 72, 128, 200, 768]))
        nq = int(rng.choice([1, 2, 127, 128, 129, 300]))
        k = int(rng.choice([1, 10, 100, 333]))
        Xd, Xr = _bf16(_unit(rng, (N, d)))
        Qd, Qr = _bf16(_unit(rng, (nq, d)))
        base = int(rng.integers(0, 10**6))
        s, i = eng.dense_topk(Xd, Qd, k, id_base=base)
        D, I = odense.flat_ip_topk(Qr, Xr, min(k + 20, N), id_base=base)
        check_topk_parity(s.cpu().numpy(), i.cpu().numpy(), D, I, k, 1e-4, what=f"fuzz-dense N={N} d={d} nq={nq} k={k}", floor=0.1)


def test_fuzz_bm25(eng):
    from legal_rag_b200.bm25_index import Bm25HostIndex
    rng = np.random.default_rng(1002)
    for trial in range(16):
        N = int(rng.choice([1, 50, 12287, 12288, 12289, 30000, 70001]))
        V = int(rng.choice([3, 40, 700]))
        p = 1.0 / np.arange(1, V + 1); p /= p.sum()
        docs = [rng.choice(V, size=int(rng.integers(1, 12)), p=p) for _ in range(N)]
        host = Bm25HostIndex.from_token_ids(docs, V)
        csr = obm25.CsrBM25.from_token_ids(docs, V)
        nq = int(rng.choice([1, 3, 40, 700]))
        queries = []
        for _ in range(min(nq, 12)):
            L = int(rng.integers(0, 10))
            q = rng.choice(V + 2, size=L).tolist()            # V, V + 1 are out of vocabulary
            queries.append([t if t < V else V + 5 for t in q])
        queries = (queries * (nq // len(queries) + 1))[:nq]
        k = int(rng.choice([1, 10, 100, 700]))
        eng.bm25_set_item_slabs(int(rng.choice([0, 1, 2])))
        try:
            qi, qt, mx = host.encode_queries(queries)
            s, i = eng.bm25_topk(host.to_device("cuda"), torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda(), max(mx, 1), k)
        finally:
            eng.bm25_set_item_slabs(0)
        m = min(len(queries), 12)
        O = [csr.search(q, min(k + 20, N)) for q in queries[:m]]
        width = max(len(o[0]) for o in O)
        check_topk_parity(s[:m].cpu().numpy(), i[:m].cpu().numpy(), np.stack([o[0] for o in O]), np.stack([o[1] for o in O]), k, 1e-3,
                          what=f"fuzz-bm25 N={N} V={V} nq={nq} k={k}")
        assert width == min(k + 20, N)


def test_fuzz_maxsim(eng):
    rng = np.random.default_rng(1003)
    for trial in range(8):
        Ld = int(rng.choice([32, 64, 128, 256]))
        Nd = int(rng.choice([1, 3, 9, 130, 1000]))
        Lq = int(rng.choice([1, 5, 32]))
        nq = int(rng.choice([1, 4, 5, 33]))
        Dd, Dr = _bf16(_unit(rng, (Nd, Ld, 128)))
        Qd, Qr = _bf16(_unit(rng, (nq, Lq, 128)))
        doclen = rng.integers(0, Ld + 1, Nd).astype(np.int32)
        allc = np.tile(np.arange(Nd, dtype=np.int64), (nq, 1))
        want = omaxsim.maxsim_scores(Qr, Dr, doclen, allc)
        got = eng.maxsim_scan_scores(Dd, torch.from_numpy(doclen).cuda(), Qd).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=3e-4, atol=3e-4, err_msg=f"scan Nd={Nd} Ld={Ld} Lq={Lq} nq={nq}")
        C = int(rng.choice([1, 17, 64]))
        cand = rng.integers(-1, Nd, (nq, C)).astype(np.int64)
        got = eng.maxsim_scores(Dd, torch.from_numpy(doclen).cuda(), Qd, torch.from_numpy(cand).cuda()).cpu().numpy()
        want = omaxsim.maxsim_scores(Qr, Dr, doclen, cand)
        ok = cand >= 0
        np.testing.assert_allclose(got[ok], want[ok], rtol=3e-4, atol=3e-4, err_msg=f"gather Nd={Nd} Ld={Ld} Lq={Lq} nq={nq} C={C}")
        assert np.isneginf(got[~ok]).all()


def test_fuzz_select_and_fuse(eng):
    rng = np.random.default_rng(1004)
    for trial in range(8):
        nq = int(rng.choice([1, 3, 9]))
        N = int(rng.choice([1, 100, 262143, 262145, 600001]))
        k = int(rng.choice([1, 50, 1024]))
        S = rng.standard_normal((nq, N)).astype(np.float32)
        S[:, :: max(1, N // 50)] = 0.5                              # tie runs
        s, i = eng.topk_select(torch.from_numpy(S).cuda(), k, id_base=7)
        D, I = odense.topk_rows(S, k, id_base=7)
        np.testing.assert_array_equal(i.cpu().numpy(), I)
        np.testing.assert_array_equal(s.cpu().numpy(), D)
    methods = ["weighted_sum", "rrf", "wrrf", "rrf_norm_blend"]
    for trial in range(12):
        kc = int(rng.choice([1, 5, 100]))
        chans = {}
        for ch in ("dense", "bm25", "colbert"):
            n = int(rng.integers(0, kc + 1))
            ids = rng.choice(400, size=n, replace=False)
            sc = np.sort(rng.uniform(-1, 30, n))[::-1]
            chans[ch] = list(zip(ids.tolist(), sc.tolist()))
        if not any(chans.values()):
            continue

        def pack(lst):
            s = torch.full((1, kc), -3.4028234663852886e38); i = torch.full((1, kc), -1, dtype=torch.int64)
            for j, (a, b) in enumerate(lst):
                s[0, j], i[0, j] = b, a
            return s.cuda(), i.cuda()
        method = methods[trial % 4]
        k = int(rng.choice([1, 10, 300]))
        fs, fi = eng.fuse_topk(pack(chans["dense"]), pack(chans["bm25"]), pack(chans["colbert"]), k=k, method=method)
        ref = ofuse.fuse(chans["dense"], chans["bm25"], chans["colbert"], method=method)
        m = min(len(ref), k + 10)
        check_topk_parity(fs.cpu().numpy(), fi.cpu().numpy(), np.array([[r["score"] for r in ref[:m]]]), np.array([[r["id"] for r in ref[:m]]]),
                          k, 1e-3, what=f"fuzz-fuse {method} kc={kc} k={k}")
