"""The C++ host front end (kasa_b200/host, binary kasa_b200/kasa_b200_cli): kASA's flags, files and outputs above
the C ABI.  CPU tests cover the host-only logic (FASTA piece/marker logic, read parsing, number formatting);
the GPU tests run `build` / `identify` on the golden inputs and compare with what the reference binary wrote
(tests/golden/*, oracle/make_golden.py): index / trie / info / frequency files byte for byte, per-read JSONL
equal after JSON parsing (every float parses to the same double, --exact-scores), profile CSV numerically.
"""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import golden_util as G
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kasa_b200", "kasa_b200_cli")


def _cli(*args, cwd=None):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kasa_b200", "csrc")])
    p = subprocess.run([CLI, *args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


def _g(case, name):
    return os.path.join(G.GOLD, case, name)


# ------------------------------------------------------------------------------------------- host only (CPU)
@pytest.mark.parametrize("case,K,six", [("k12", 12, False), ("k12_build6", 12, True), ("k25", 25, False), ("k12", 12, True)])
def test_build_strings_match_oracle(case, K, six):
    """readFasta piece / marker logic incl. the 2048-byte chunk quirk (Read.hpp:2306-2578), reading .gz."""
    args = ["host-build-strings", "-c", _g(case, "content.txt"), "-i", _g(case, "ref.fa.gz"), "--kH", str(K)] + (["--six"] if six else [])
    rc, out, err = _cli(*args)
    assert rc == 0, err
    got = [(l.split(b"\t")[1], int(l.split(b"\t")[0])) for l in out.split(b"\n") if l]
    want = [(s, t) for s, t in O.fasta_build_strings(G.gz(case, "ref.fa.gz"), G.acc2tax(case), K, six) if s]
    assert got == want


def test_read_parsing_fastq_fasta_gz(tmp_path):
    rc, out, err = _cli("host-reads", "-i", _g("k12", "reads.fq.gz"))
    assert rc == 0, err
    got = [l.split(b"\t") for l in out.split(b"\n") if l]
    want = G.fastq("k12")
    assert [(g[0], g[2]) for g in got] == [(n, s) for n, s in want if True]
    assert all(g[1] == b"1" for g in got)
    # multi-line FASTA, CRLF, lower case and an empty record
    fa = tmp_path / "r.fa"
    fa.write_bytes(b">a desc\r\nACGT\r\nacgtn\r\n>b\n>c\nTTTT\nGG\nA\n")
    rc, out, err = _cli("host-reads", "-i", str(fa))
    assert rc == 0, err
    got = [l.split(b"\t") for l in out.split(b"\n") if l]
    assert [(g[0], g[1], g[2]) for g in got] == [(b"a desc", b"2", b"ACGTacgtn"), (b"b", b"0", b""), (b"c", b"3", b"TTTTGGA")]


def test_cli_errors():
    rc, _, err = _cli("identify", "-c", "/nonexistent", "-d", "/nonexistent", "-i", "/nonexistent")
    assert rc == 1 and b"ERROR" in err
    rc, _, err = _cli("shrink")
    assert rc == 1 and b"not supported" in err


# ------------------------------------------------------------------------------------------- through the GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case,flags", [("k12", []), ("k12_build6", ["--six"]), ("k25", ["--kH", "25"])])
def test_cli_build_files_byte_identical(case, flags, tmp_path):
    idx = str(tmp_path / "idx")
    rc, out, err = _cli("build", "-c", _g(case, "content.txt"), "-d", idx, "-i", _g(case, "ref.fa.gz"), "-m", "4", "-n", "1", *flags)
    assert rc == 0, err
    info = open(idx + "_info.txt").read()
    assert info == G.text(case, "idx_info.txt")
    n = int(info.split()[0])
    rec = 20 if "128" in info.split()[1:] else 12
    data = open(idx, "rb").read()
    assert data[:n * rec] == G.gz(case, "idx.gz")
    assert len(data) % (2048000 if rec == 20 else 2101248) == 0 and not any(data[n * rec:])
    assert open(idx + "_trie.txt").read() == G.text(case, "idx_trie.txt")
    nt = int(G.text(case, "idx_trie.txt").split()[0])
    assert open(idx + "_trie", "rb").read()[:nt * 12] == G.gz(case, "idx_trie.gz")
    assert open(idx + "_f.txt").read() == G.text(case, "idx_f.txt")


@pytest.mark.gpu
def test_cli_build_with_another_codon_table(tmp_path):
    """`-a <gc.prt> <id>` (kASA::setCodonTable, kASA.hpp:579-615): the index equals the one `kASA build -a` wrote for the same file
    (translation table 4: TGA = W instead of stop; tests/golden/k12/alt_codon_sha256.txt from oracle/_ref/kASA_scalar)."""
    import hashlib
    idx = str(tmp_path / "idx")
    rc, out, err = _cli("build", "-c", _g("k12", "content.txt"), "-d", idx, "-i", _g("k12", "ref.fa.gz"), "-a", _g("k12", "gc.prt"), "4")
    assert rc == 0, err
    n_want, sha_want = G.text("k12", "alt_codon_sha256.txt").split()
    n = int(open(idx + "_info.txt").read().split()[0])
    assert n == int(n_want)
    data = open(idx, "rb").read()[:n * 12]
    assert hashlib.sha256(data).hexdigest() == sha_want
    assert data != G.gz("k12", "idx.gz")          # and it differs from the standard-code index


@pytest.mark.gpu
def test_cli_build_one_frame(tmp_path):
    """`build --one`: the index file equals the one `kASA build --one` wrote (tests/golden/k12/one_build_sha256.txt)."""
    import hashlib
    idx = str(tmp_path / "idx")
    rc, out, err = _cli("build", "-c", _g("k12", "content.txt"), "-d", idx, "-i", _g("k12", "ref.fa.gz"), "--one")
    assert rc == 0, err
    n_want, sha_want = G.text("k12", "one_build_sha256.txt").split()[:2]
    n = int(open(idx + "_info.txt").read().split()[0])
    assert n == int(n_want) and hashlib.sha256(open(idx, "rb").read()[:n * 12]).hexdigest() == sha_want


def _write_index(case, tmp_path):
    """golden index files laid out the way `identify -d` expects them."""
    idx = str(tmp_path / "idx")
    open(idx, "wb").write(G.gz(case, "idx.gz"))
    for suf in ("_info.txt", "_f.txt", "_trie.txt"):
        open(idx + suf, "w").write(G.text(case, "idx" + suf))
    open(idx + "_trie", "wb").write(G.gz(case, "idx_trie.gz"))
    return idx


def _check_profile(path, case, tag):
    want = G.text(case, tag + ".csv").splitlines()
    got = open(path).read().splitlines()
    assert got[0] == want[0]
    assert [l.split(",")[:2] for l in got] == [l.split(",")[:2] for l in want]      # same taxa, same order
    for lg, lw in zip(got[1:], want[1:]):
        a = np.array(lg.split(",")[2:], dtype=np.float64)
        b = np.array(lw.split(",")[2:], dtype=np.float64)
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-12)                       # printed with 6 significant digits


@pytest.mark.gpu
@pytest.mark.parametrize("case,tag,flags", [
    ("k12", "six_12_7", ["--six"]),
    ("k12", "three_11_8", ["-k", "11", "8"]),
    ("k12", "six_12_12", ["--six", "-k", "12", "12"]),
    ("k25", "six_25_7", ["--six", "-k", "25", "7"]),
    ("k12", "one_12_7", ["--one"]),
])
def test_cli_identify_matches_reference_outputs(case, tag, flags, tmp_path):
    idx = _write_index(case, tmp_path)
    inp = ["-1", _g(case, "reads_1.fq.gz"), "-2", _g(case, "reads_2.fq.gz")] if case == "k25" else ["-i", _g(case, "reads.fq.gz")]
    prof, rtt = str(tmp_path / "p.csv"), str(tmp_path / "q.jsonl")
    rc, out, err = _cli("identify", "-c", _g(case, "content.txt"), "-d", idx, *inp, "-p", prof, "-q", rtt, "--jsonl", "-b", "100000",
                        "-v", "--exact-scores", *flags)      # one batch, like the reference run: its fp32 order depends on the batch
    assert rc == 0, err
    assert ("Number of k-mers in input: %d " % G.n_kmers(case, tag)).encode() in out
    _check_profile(prof, case, tag)
    want = [json.loads(l) for l in G.gz(case, tag + ".jsonl.gz").decode().splitlines()]
    got = [json.loads(l) for l in open(rtt).read().splitlines()]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g == w, (g, w)


@pytest.mark.gpu
def test_cli_batching_does_not_change_output(tmp_path):
    """default scores (float32 sums in ascending k-mer order per read) are independent of how the reads are cut into batches."""
    case = "k25"
    idx = _write_index(case, tmp_path)
    base = ["identify", "-c", _g(case, "content.txt"), "-d", idx, "-1", _g(case, "reads_1.fq.gz"), "-2", _g(case, "reads_2.fq.gz"),
            "--six", "-k", "25", "7", "--jsonl", "-b", "100000"]
    outs = []
    # the input is streamed: a batch closes at --batch-reads reads or --batch-bases bases (2000 bases = 6 pairs of 2 x 150 bp)
    for i, extra in enumerate([["--batch-reads", "7"], ["--batch-reads", "64"], ["--batch-reads", "100000"], ["--batch-bases", "2000"]]):
        rc, _, err = _cli(*base, "-p", str(tmp_path / ("p%d" % i)), "-q", str(tmp_path / ("q%d" % i)), *extra)
        assert rc == 0, err
        outs.append((open(tmp_path / ("p%d" % i)).read(), open(tmp_path / ("q%d" % i)).read()))
    assert outs[0] == outs[1] == outs[2] == outs[3]
    _check_profile(str(tmp_path / "p0"), case, "six_25_7")


@pytest.mark.gpu
def test_cli_multi_gpu_round_robin_keeps_output_and_order(tmp_path, monkeypatch):
    """--gpus N deals the batches round-robin over N contexts and writes the reads in input order; here the N contexts
    share one device (KASA_CLI_SAME_DEVICE), which exercises the same host logic a multi-GPU box runs."""
    case = "k12"
    idx = _write_index(case, tmp_path)
    base = ["identify", "-c", _g(case, "content.txt"), "-d", idx, "-i", _g(case, "reads.fq.gz"), "--six", "--jsonl", "-b", "100000"]
    monkeypatch.setenv("KASA_CLI_SAME_DEVICE", "1")
    outs = []
    for i, (gpus, br) in enumerate([("1", "100000"), ("3", "17"), ("2", "50"), ("4", "1000")]):
        rc, _, err = _cli(*base, "-p", str(tmp_path / ("p%d" % i)), "-q", str(tmp_path / ("q%d" % i)), "--batch-reads", br, "--gpus", gpus)
        assert rc == 0, err
        outs.append(open(tmp_path / ("q%d" % i)).read())
        _check_profile(str(tmp_path / ("p%d" % i)), case, "six_12_7")
    assert outs[0] == outs[1] == outs[2] == outs[3]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["json", "tsv", "kraken"])
def test_cli_other_formats_consistent_with_jsonl(fmt, tmp_path):
    """default -b 3: the json / tsv / kraken writers list the same taxa and scores as the jsonl writer."""
    case, tag = "k12", "six_12_7"
    idx = _write_index(case, tmp_path)
    base = ["identify", "-c", _g(case, "content.txt"), "-d", idx, "-i", _g(case, "reads.fq.gz"), "--six", "--exact-scores"]
    rc, _, err = _cli(*base, "-q", str(tmp_path / "a.jsonl"), "--jsonl")
    assert rc == 0, err
    rc, _, err = _cli(*base, "-q", str(tmp_path / "b.out"), "--" + fmt)
    assert rc == 0, err
    ref = [json.loads(l) for l in open(tmp_path / "a.jsonl").read().splitlines()]
    text = open(tmp_path / "b.out").read()
    if fmt == "json":
        got = json.loads(text)
        assert got == ref
    elif fmt == "tsv":
        rows = [l.split("\t") for l in text.splitlines()[1:]]
        assert len(rows) == len(ref)
        full_list = [json.loads(l) for l in G.gz(case, tag + ".jsonl.gz").decode().splitlines()]     # reference, -b 100000
        for r, j, fj in zip(rows, ref, full_list):
            hits = j["Top hits"] + j["Further hits"]
            assert int(r[0]) == j["Read number"] and r[1] == j["Specifier from input file"]
            if not hits:
                assert r[2:] == ["-", "-", "-", "-"]
                continue
            # the tsv writer counts distinct k-mer scores from the first entry (Compare.hpp:1622-1640), so it can list
            # more taxa than "Top hits" + "Further hits"; both are prefixes of the same sorted list
            full = [h for h in (fj["Top hits"] + fj["Further hits"])]
            want, distinct, before = [], 0, 0.0
            for h in full:
                if distinct >= 3:
                    break
                want.append(h)
                if h["k-mer Score"] != before:
                    before, distinct = h["k-mer Score"], distinct + 1
            assert r[2].split(";") == [h["tax ID"] for h in want]
            assert r[3].split(";") == [h["Name"] for h in want]
            assert [float(x.split(",")[1]) for x in r[4].split(";")] == [h["k-mer Score"] for h in want]
            assert [float(x.split(",")[0]) for x in r[4].split(";")] == [h["Relative Score"] for h in want]
            assert [float(x) for x in r[5].split(";")] == [h["Error"] for h in want]
    else:
        rows = [l.split("\t") for l in text.splitlines()]
        assert len(rows) == len(ref)
        for r, j in zip(rows, ref):
            hits = j["Top hits"] + j["Further hits"]
            assert r[0] == ("C" if hits else "U")
            if hits:
                assert r[2] == hits[0]["tax ID"] and int(r[3]) == j["Length"]
                assert [x.split(":")[0] for x in r[4].split()] == [h["tax ID"] for h in hits]


# ---------------------------------------------------------------------------------------------------------------------------------
# index-maintenance modes: trie / getFrequency / delete / merge against the reference binary's files
# (tests/golden/k12/modes_sha256.txt and modes/, written by oracle/make_golden_modes.py from kASA_scalar)
def _modes():
    want = {}
    for line in G.text("k12", "modes_sha256.txt").splitlines():
        if line and not line.startswith("#"):
            name, nbytes, sha = line.split()
            want[name] = (int(nbytes), sha)
    return want


def _records_sha(path, rec=12):
    import hashlib
    n = int(open(path + "_info.txt").read().split()[0])
    return n * rec, hashlib.sha256(open(path, "rb").read()[:n * rec]).hexdigest()


def _trie_sha(path):
    import hashlib
    n = int(open(path + "_trie.txt").read().split()[0])
    return n * 12, hashlib.sha256(open(path + "_trie", "rb").read()[:n * 12]).hexdigest()


@pytest.mark.gpu
def test_cli_trie_and_frequency_modes_rewrite_the_reference_files(tmp_path):
    """`trie -d idx` and `getFrequency -c content -d idx` from the index file alone (main.cpp:1422-1458, 1336-1362)."""
    import hashlib
    want = _modes()
    idx = str(tmp_path / "idx")
    open(idx, "wb").write(G.gz("k12", "idx.gz"))
    open(idx + "_info.txt", "w").write(G.text("k12", "idx_info.txt"))
    rc, out, err = _cli("trie", "-d", idx)
    assert rc == 0, err
    assert _trie_sha(idx) == want["trie/t_trie"]
    rc, out, err = _cli("getFrequency", "-c", _g("k12", "content.txt"), "-d", idx)
    assert rc == 0, err
    data = open(idx + "_f.txt", "rb").read()
    assert (len(data), hashlib.sha256(data).hexdigest()) == want["getFrequency/f_f.txt"]
    assert data.decode() == G.text("k12", "idx_f.txt")


@pytest.mark.gpu
def test_cli_delete_mode_matches_the_reference(tmp_path):
    """`delete -l delnodes.dmp`: index, prefix table and frequency file equal the ones `kASA delete` wrote (Update.hpp:28-94)."""
    want = _modes()
    idx = _write_index("k12", tmp_path)
    out_idx = str(tmp_path / "del")
    rc, out, err = _cli("delete", "-c", _g("k12", "content.txt"), "-d", idx, "-o", out_idx, "-l", _g("k12", os.path.join("modes", "del.dmp")))
    assert rc == 0, err
    assert _records_sha(out_idx) == want["delete/del"]
    assert _trie_sha(out_idx) == want["delete/del_trie"]
    assert open(out_idx + "_f.txt").read() == G.text("k12", os.path.join("modes", "del_f.txt"))


@pytest.mark.gpu
def test_cli_merge_mode_matches_the_reference(tmp_path):
    """Two indexes built from the halves of ref.fa, merged: the halves and the merged index + prefix table equal the reference's
    (`kASA merge`, Read.hpp:3180-3243), which equal the one-shot build.  The reference's merge writes no _info.txt and therefore a
    frequency file of zeros (modes/m_f.txt); ours writes both, the frequencies being those of the one-shot build."""
    import gzip
    want = _modes()
    text = gzip.open(_g("k12", "ref.fa.gz"), "rt").read()
    recs = [">" + r for r in text.split(">") if r]
    h = len(recs) // 2
    content = _g("k12", "content.txt")
    for name, part in (("a", recs[:h]), ("b", recs[h:])):
        open(tmp_path / (name + ".fa"), "w").write("".join(part))
        rc, out, err = _cli("build", "-c", content, "-d", str(tmp_path / name), "-i", str(tmp_path / (name + ".fa")))
        assert rc == 0, err
        assert _records_sha(str(tmp_path / name)) == want["merge/" + name]
    m = str(tmp_path / "m")
    rc, out, err = _cli("merge", "-c1", content, "-c2", content, "-co", m + "_content.txt", "--firstIndex", str(tmp_path / "a"),
                        "--secondIndex", str(tmp_path / "b"), "-o", m)
    assert rc == 0, err
    assert _records_sha(m) == want["merge/m"]
    assert _trie_sha(m) == want["merge/m_trie"]
    assert open(m + "_content.txt").read() == open(content).read()
    assert open(m + "_f.txt").read() == G.text("k12", "idx_f.txt")
    assert all(f == "0" for l in G.text("k12", os.path.join("modes", "m_f.txt")).splitlines() for f in l.split("\t")[1:])     # the reference's file
    rc, out, err = _cli("merge", "-c", content, "--firstIndex", m, "--secondIndex", m, "-o", str(tmp_path / "mm"))
    assert rc != 0 and b"different" in err


@pytest.mark.gpu
def test_cli_update_mode_matches_the_reference(tmp_path):
    """`update -c content -d a -i b.fa -o u`: the first half's index plus the second half's sequences = the files `kASA update`
    wrote (Update.hpp:99-), which equal the one-shot build."""
    import gzip
    want = _modes()
    text = gzip.open(_g("k12", "ref.fa.gz"), "rt").read()
    recs = [">" + r for r in text.split(">") if r]
    h = len(recs) // 2
    content = _g("k12", "content.txt")
    open(tmp_path / "a.fa", "w").write("".join(recs[:h]))
    open(tmp_path / "b.fa", "w").write("".join(recs[h:]))
    a, u = str(tmp_path / "a"), str(tmp_path / "u")
    rc, out, err = _cli("build", "-c", content, "-d", a, "-i", str(tmp_path / "a.fa"))
    assert rc == 0, err
    rc, out, err = _cli("update", "-c", content, "-d", a, "-i", str(tmp_path / "b.fa"), "-o", u)
    assert rc == 0, err
    assert _records_sha(u) == want["update/u"]
    assert _trie_sha(u) == want["update/u_trie"]
    assert open(u + "_f.txt").read() == G.text("k12", "idx_f.txt")
    # without -o the index is updated in place
    rc, out, err = _cli("update", "-c", content, "-d", a, "-i", str(tmp_path / "b.fa"))
    assert rc == 0, err
    assert _records_sha(a) == want["update/u"]


@pytest.mark.gpu
def test_cli_identify_multiple_names_and_equals_identify(tmp_path):
    """identify_multiple (main.cpp:1118-1334): every file of a folder against the index loaded once.  The output names are the ones the
    reference binary produced (modes_sha256.txt: <prefix><file name without its last extension><ending>), the contents those of
    `identify` on the same file — which the six_12_7 goldens pin against the reference."""
    import gzip
    names = sorted(k.split("/", 1)[1] for k in _modes() if k.startswith("identify_multiple/"))
    assert names == ["p_first.part.csv", "p_second.csv", "q_first.part.jsonl", "q_second.jsonl"]
    idx = _write_index("k12", tmp_path)
    fq = gzip.open(_g("k12", "reads.fq.gz"), "rt").read().splitlines(keepends=True)
    half = (len(fq) // 8) * 4
    os.makedirs(tmp_path / "in")
    os.makedirs(tmp_path / "res")
    open(tmp_path / "in" / "first.part.fq", "w").write("".join(fq[:half]))
    open(tmp_path / "in" / "second.fq", "w").write("".join(fq[half:]))
    common = ["-c", _g("k12", "content.txt"), "-d", idx, "--jsonl", "--six", "--batch-reads", "64"]
    rc, out, err = _cli("identify_multiple", *common, "-i", str(tmp_path / "in") + "/", "-q", str(tmp_path / "res" / "q_"), "-p", str(tmp_path / "res" / "p_"))
    assert rc == 0, err
    assert sorted(os.listdir(tmp_path / "res")) == names
    for part, stem in (("first.part.fq", "first.part"), ("second.fq", "second")):
        rc, out, err = _cli("identify", *common, "-i", str(tmp_path / "in" / part), "-q", str(tmp_path / "single.jsonl"), "-p", str(tmp_path / "single.csv"))
        assert rc == 0, err
        assert open(tmp_path / "single.jsonl").read() == open(tmp_path / "res" / ("q_" + stem + ".jsonl")).read()
        assert open(tmp_path / "single.csv").read() == open(tmp_path / "res" / ("p_" + stem + ".csv")).read()


@pytest.mark.gpu
def test_cli_identify_with_a_halved_index(tmp_path):
    """An index halved by `kASA half` (info file "<n>\\n3", 6-byte records, the upper letters in _trie) gives the output of the full index —
    as it does with the reference binary (modes_sha256.txt: half/identify_equals_full_index)."""
    _, taxids, _ = G.content("k12")
    rec6, trie = O.half_index(G.index("k12"), taxids)
    full = _write_index("k12", tmp_path)
    h = str(tmp_path / "idxh")
    rec6.tofile(h)
    trie.tofile(h + "_trie")
    open(h + "_info.txt", "w").write("%d\n3" % len(rec6))
    open(h + "_trie.txt", "w").write(str(len(trie)))
    open(h + "_f.txt", "w").write(G.text("k12", "idx_f.txt"))
    outs = []
    for idx, tag in ((full, "full"), (h, "half")):
        rc, out, err = _cli("identify", "-c", _g("k12", "content.txt"), "-d", idx, "-i", _g("k12", "reads.fq.gz"), "--jsonl", "--six",
                            "-q", str(tmp_path / (tag + ".jsonl")), "-p", str(tmp_path / (tag + ".csv")))
        assert rc == 0, err
        outs.append((open(tmp_path / (tag + ".jsonl")).read(), open(tmp_path / (tag + ".csv")).read()))
    assert outs[0] == outs[1]


def test_cli_half_mode_matches_the_reference(tmp_path):
    """`half` (main.cpp:1465-1474): a streaming pass on the host, no device needed.  Records, prefix table and info file equal the ones
    `kASA half` wrote (modes_sha256.txt); the frequency file is the old one."""
    want = _modes()
    idx = _write_index("k12", tmp_path)
    h = str(tmp_path / "idxh")
    rc, out, err = _cli("half", "-c", _g("k12", "content.txt"), "-d", idx, "-o", h)
    assert rc == 0, err
    info = open(h + "_info.txt").read().split()
    assert info[1] == "3"
    assert _records_sha(h, rec=6) == want["half/idxh"]
    assert _trie_sha(h) == want["half/idxh_trie"]
    assert open(h + "_f.txt").read() == G.text("k12", "idx_f.txt")
    rc, out, err = _cli("half", "-c", _g("k12", "content.txt"), "-d", idx, "-o", idx)
    assert rc != 0 and b"same" in err
    # the index cut after a prefix run of four records (the last run is written one short, as the reference writes it)
    cut = [int(k.split("_")[-2]) for k in want if k.startswith("half/cuth_of_the_first_")][0]
    c = str(tmp_path / "cut")
    open(c, "wb").write(G.gz("k12", "idx.gz")[:cut * 12])
    open(c + "_info.txt", "w").write(str(cut))
    open(c + "_f.txt", "w").write(G.text("k12", "idx_f.txt"))
    rc, out, err = _cli("half", "-c", _g("k12", "content.txt"), "-d", c, "-o", c + "h")
    assert rc == 0, err
    assert _records_sha(c + "h", rec=6) == want["half/cuth_of_the_first_%d_records" % cut] and _trie_sha(c + "h") == want["half/cuth_trie"]


def _yaml(path, **kv):
    """kASA's parameters.yaml (reference repo root): sections with flat `Key: value` lines, strings quoted."""
    base = {"Mode": '"build"', "Index": '""', "ContentFile": '""', "kHigh": 12, "kLow": 7, "NumberOfThreads": 2, "AvailableRAMinGB": 5,
            "FilePathForTemporaryFiles": '""', "CallIndex": 0, "Verbose": "false", "AlphabetFile": '""', "AlphabetIndex": 0,
            "InputFileOrFolder": '""', "PairedEnd-First": '""', "PairedEnd-Second": '""', "AlreadyTranslated": "false",
            "One": "false", "Three": "true", "Six": "false", "ProfileOutputfile": '""', "ReadIDtoTaxIDOutputfile": '""',
            "ReadIDtoTaxIDOutputFormat": '"json"', "UseRAMOnly": "false", "NumberOfTaxaPerRead": 3, "UniqueKmersOnly": "false",
            "ThresholdForScore": 0.0, "PrintCoverage": "false", "Filter": '"_ _"', "FileWithDeletedTaxa": '""', "ShrinkingStrategy": 2,
            "FirstOldIndex": '""', "SecondOldIndex": '""', "NewIndex": '""'}
    base.update(kv)
    with open(path, "w") as f:
        f.write("# config file for kASA\n\nGeneral: \n # modes available: build, identify, ...\n")
        for k, v in base.items():
            f.write(" %s: %s\n" % (k, v))
    return str(path)


def test_cli_parameters_file_is_read_like_the_flags(tmp_path):
    """`--parameters <yaml>` (main.cpp:264, Utilities.hpp:1114-): the keys of kASA's parameters.yaml land where the flags land."""
    y = _yaml(tmp_path / "p.yaml", Mode='"identify"', Index='"/data/idx"', ContentFile='"/data/content.txt"', kHigh=11, kLow=8, Verbose="true",
              InputFileOrFolder='"/data/reads.fq"', Three="false", Six="true", ProfileOutputfile='"/out/p.csv"', ReadIDtoTaxIDOutputfile='"/out/q.jsonl"',
              ReadIDtoTaxIDOutputFormat='"jsonl"', NumberOfTaxaPerRead=5, ThresholdForScore=0.25)
    rc, out, err = _cli("--parameters", y, "--show")
    assert rc == 0, err
    got = dict(l.split("=", 1) for l in out.decode().splitlines() if "=" in l)
    assert got["mode"] == "identify" and got["index"] == "/data/idx" and got["content"] == "/data/content.txt" and got["input"] == "/data/reads.fq"
    assert got["k"] == "11 8" and got["frames"] == "6" and got["format"] == "jsonl" and got["beasts"] == "5" and got["threshold"] == "0.25"
    assert got["profile"] == "/out/p.csv" and got["rtt"] == "/out/q.jsonl" and got["verbose"] == "1" and got["highest_k"] == "12"
    # features this front end does not have are refused when switched on; two frames at once as well
    rc, out, err = _cli("--parameters", _yaml(tmp_path / "q.yaml", PrintCoverage="true"), "--show")
    assert rc != 0 and b"PrintCoverage" in err
    rc, out, err = _cli("--parameters", _yaml(tmp_path / "r.yaml", Six="true"), "--show")
    assert rc != 0 and b"frames" in err
    rc, out, err = _cli("--parameters", _yaml(tmp_path / "s.yaml", kHigh=25, Mode='"build"'), "--show")
    assert rc == 0 and b"highest_k=25" in out


def test_cli_parameters_file_drives_a_mode(tmp_path):
    """the `half` mode run from a parameter file (host only): the same files as with flags."""
    want = _modes()
    idx = _write_index("k12", tmp_path)
    h = str(tmp_path / "idxh")
    y = _yaml(tmp_path / "p.yaml", Mode='"half"', Index='"%s"' % idx, ContentFile='"%s"' % _g("k12", "content.txt"), NewIndex='"%s"' % h)
    rc, out, err = _cli("--parameters", y)
    assert rc == 0, err
    assert _records_sha(h, rec=6) == want["half/idxh"] and _trie_sha(h) == want["half/idxh_trie"]


def test_cli_redundancy_mode_prints_what_the_reference_prints(tmp_path):
    """`redundancy -v` (main.cpp:1364-1420, Shrink::histogram): histogram of taxa per k-mer and the advice line, host only."""
    idx = _write_index("k12", tmp_path)
    rc, out, err = _cli("redundancy", "-c", _g("k12", "content.txt"), "-d", idx, "-v")
    assert rc == 0, err
    got = [l for l in out.decode().splitlines() if not l.startswith("OUT: Time")]
    assert got == G.text("k12", os.path.join("modes", "redundancy.txt")).splitlines()
