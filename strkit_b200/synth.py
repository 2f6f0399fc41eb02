"""Synthetic HiFi-/ONT-like read sets at synthetic loci (the BASELINE.json configs).

Emits exactly the tuple the hot path sees -- (start-count estimate, tr_seq_wc, flank_left_seq_wc[-70:],
flank_right_seq_wc[:70], motif) per read, grouped by locus (call_locus.py:1129-1155) -- already packed as
a ReadBatch.  The bulk generator is vectorised torch (CPU or CUDA tensors; torch is used here for tensor
staging only), so a 30-million-read set never passes through per-read Python.

Generator definition: motif = iid ACGT of length m, rejected if it has a proper
sub-period; k_ref copies; 70-nt iid flanks rejected if the m bases next to the tract equal the motif;
diploid alleles k_ref + delta; read = fl + motif*a + fr through an error channel (substitution /
insertion / deletion / stutter of one unit / low-quality 'X' wildcards); est_cn = round(len(tr) / m).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .batcher import LocusReads, ReadBatch, pack_loci

__all__ = ["SynthSpec", "CONFIGS", "generate", "generate_expansions", "SynthBatch"]

_ASCII = torch.tensor(list(b"ACGTRYSWKMBDHVNX"), dtype=torch.uint8)
FLANK = 70


@dataclass(frozen=True)
class SynthSpec:
    name: str
    motif_min: int = 2
    motif_max: int = 6
    k_min: int = 10
    k_max: int = 60
    reads_per_locus: int = 30
    sub: float = 0.0005
    ins: float = 0.001
    dele: float = 0.001
    stutter: float = 0.02
    x_rate: float = 0.001
    delta_p: tuple = (0.6, 0.25, 0.1, 0.05)  # P(delta = 0), P(|delta| = 1), P(|delta| = 2), P(|delta| in 3..5)
    config_id: int = 1


CONFIGS = {
    1: SynthSpec("cfg1-hifi-1k-loci", config_id=1),
    2: SynthSpec("cfg2-hifi-genome-wide", config_id=2),
    3: SynthSpec("cfg3-ont-noisy", reads_per_locus=40, sub=0.02, ins=0.015, dele=0.015, x_rate=0.01,
                 delta_p=(0.4, 0.3, 0.15, 0.15), config_id=3),
    5: SynthSpec("cfg5-full-pipeline", config_id=5),
}


@dataclass
class SynthBatch:
    """Torch tensors (on the generating device) + conversion to the host ReadBatch."""
    arena: torch.Tensor
    seq_off: torch.Tensor
    lens: torch.Tensor
    est_cn: torch.Tensor
    read_begin: torch.Tensor
    motif_off: torch.Tensor
    motif_len: torch.Tensor
    true_cn: torch.Tensor = field(default=None)  # copies actually written into each read (before errors)

    def to_host(self, pin: bool = False, nibble: bool = False) -> ReadBatch:
        """nibble=True: the arena goes out nibble-packed (batcher.ARENA_NIBBLE), packed on the generating device."""
        def h(t, dt):
            a = t.detach().to("cpu")
            if pin:
                a = a.pin_memory()
            return a.numpy().view(dt) if dt is not None else a.numpy()

        arena = self.arena
        if nibble:
            lut = torch.full((256,), 16, dtype=torch.uint8, device=arena.device)
            lut[_ASCII.to(arena.device).long()] = torch.arange(16, dtype=torch.uint8, device=arena.device)
            codes = lut[arena.long()]
            if codes.numel() & 1:
                codes = torch.cat([codes, codes.new_zeros(1)])
            arena = codes[0::2] | (codes[1::2] << 4)
        return ReadBatch(arena=h(arena, None), seq_off=h(self.seq_off, np.uint64), lens=h(self.lens, None),
                         est_cn=h(self.est_cn, None), read_begin=h(self.read_begin, None),
                         motif_off=h(self.motif_off, np.uint64), motif_len=h(self.motif_len, None),
                         arena_format=1 if nibble else 0)


def _rand_bases(shape, gen, dev):
    return torch.randint(0, 4, shape, generator=gen, device=dev, dtype=torch.int64)


def _sub_periodic(motif: torch.Tensor, m: torch.Tensor, mmax: int) -> torch.Tensor:
    """True where the motif equals itself shifted by a proper divisor period (incl. homopolymers)."""
    nl = motif.shape[0]
    idx = torch.arange(mmax, device=motif.device).expand(nl, mmax)
    inside = idx < m[:, None]
    bad = torch.zeros(nl, dtype=torch.bool, device=motif.device)
    for p in range(1, mmax):
        divides = (m % p == 0) & (m > p)
        same = ((motif == torch.gather(motif, 1, idx % p)) | ~inside).all(dim=1)
        bad |= divides & same
    return bad


def generate(spec: SynthSpec, n_loci: int, seed: int | None = None, device: str = "cpu",
             chunk_loci: int = 4096) -> SynthBatch:
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + 1000 * spec.config_id if seed is None else seed)
    parts = []
    for lo in range(0, n_loci, chunk_loci):
        parts.append(_generate_chunk(spec, min(chunk_loci, n_loci - lo), gen, dev))
    # concatenate chunks: arena = [all reads][all motifs]
    read_bytes = [p["reads"].numel() for p in parts]
    total_read_bytes = sum(read_bytes)
    seq_off, motif_off, read_begin = [], [], []
    rb, mb, nr = 0, 0, 0
    for p in parts:
        seq_off.append(p["seq_off"] + rb)
        motif_off.append(p["motif_off"] + total_read_bytes + mb)
        read_begin.append(p["read_begin"][:-1] + nr)
        rb += p["reads"].numel()
        mb += p["motifs"].numel()
        nr += p["est_cn"].numel()
    read_begin.append(torch.tensor([nr], dtype=torch.int64, device=dev))
    return SynthBatch(
        arena=torch.cat([p["reads"] for p in parts] + [p["motifs"] for p in parts]),
        seq_off=torch.cat(seq_off), lens=torch.cat([p["lens"] for p in parts]),
        est_cn=torch.cat([p["est_cn"] for p in parts]), read_begin=torch.cat(read_begin),
        motif_off=torch.cat(motif_off), motif_len=torch.cat([p["motif_len"] for p in parts]),
        true_cn=torch.cat([p["true_cn"] for p in parts]))


def _generate_chunk(spec: SynthSpec, nl: int, gen, dev):
    mmax = spec.motif_max
    rpl = spec.reads_per_locus
    m = torch.randint(spec.motif_min, spec.motif_max + 1, (nl,), generator=gen, device=dev)
    motif = _rand_bases((nl, mmax), gen, dev)
    for _ in range(64):
        bad = _sub_periodic(motif, m, mmax)
        if not bool(bad.any()):
            break
        motif[bad] = _rand_bases((int(bad.sum()), mmax), gen, dev)
    k_ref = torch.randint(spec.k_min, spec.k_max + 1, (nl,), generator=gen, device=dev)
    fl = _rand_bases((nl, FLANK), gen, dev)
    fr = _rand_bases((nl, FLANK), gen, dev)
    idx = torch.arange(mmax, device=dev).expand(nl, mmax)
    inside = idx < m[:, None]
    for _ in range(64):
        # last m bases of the left flank == motif, or first m bases of the right flank == motif
        fl_tail = torch.gather(fl, 1, (FLANK - m[:, None] + idx).clamp(0, FLANK - 1))
        bad_l = ((fl_tail == motif) | ~inside).all(dim=1)
        bad_r = ((fr[:, :mmax] == motif) | ~inside).all(dim=1)
        if not bool((bad_l | bad_r).any()):
            break
        fl[bad_l] = _rand_bases((int(bad_l.sum()), FLANK), gen, dev)
        fr[bad_r] = _rand_bases((int(bad_r.sum()), FLANK), gen, dev)

    # diploid alleles
    p0, p1, p2, p3 = spec.delta_p
    probs = torch.tensor([p3 / 6] * 3 + [p2 / 2, p1 / 2, p0, p1 / 2, p2 / 2] + [p3 / 6] * 3, device=dev)
    delta = torch.multinomial(probs, nl * 2, replacement=True, generator=gen).view(nl, 2) - 5
    alleles = (k_ref[:, None] + delta).clamp(min=1)
    nr = nl * rpl
    locus = torch.arange(nl, device=dev).repeat_interleave(rpl)
    pick = torch.randint(0, 2, (nr,), generator=gen, device=dev)
    a = alleles[locus, pick]
    u = torch.rand(nr, generator=gen, device=dev)
    a = (a + (u < spec.stutter / 2).long() - ((u >= spec.stutter / 2) & (u < spec.stutter)).long()).clamp(min=1)

    mr = m[locus]
    T = mr * a
    L = FLANK + T + FLANK
    starts = torch.cumsum(L, 0) - L
    total = int(L.sum())
    read_id = torch.arange(nr, device=dev).repeat_interleave(L)
    pos = torch.arange(total, device=dev) - starts[read_id]
    Tr = T[read_id]
    loc = locus[read_id]
    region = (pos >= FLANK).long() + (pos >= FLANK + Tr).long()
    base = torch.where(
        region == 0, fl[loc, pos.clamp(max=FLANK - 1)],
        torch.where(region == 1, motif[loc, (pos - FLANK).clamp(min=0) % mr[read_id]],
                    fr[loc, (pos - FLANK - Tr).clamp(0, FLANK - 1)]))
    del Tr, loc

    # error channel
    u = torch.rand(total, generator=gen, device=dev)
    keep = u >= spec.dele
    is_sub = keep & (u < spec.dele + spec.sub)
    base = torch.where(is_sub, (base + torch.randint(1, 4, (total,), generator=gen, device=dev)) % 4, base)
    has_ins = torch.rand(total, generator=gen, device=dev) < spec.ins
    counts = keep.long() + has_ins.long()
    out_total = int(counts.sum())
    src = torch.arange(total, device=dev).repeat_interleave(counts)
    grp_start = torch.cumsum(counts, 0) - counts
    k_in = torch.arange(out_total, device=dev) - grp_start[src]
    inserted = (k_in == 1) | ((k_in == 0) & ~keep[src])
    out_base = torch.where(inserted, _rand_bases((out_total,), gen, dev), base[src])
    out_region = region[src]
    out_read = read_id[src]
    del src, grp_start, k_in, inserted, base, region, read_id, pos, u, keep, is_sub, has_ins, counts
    if spec.x_rate > 0:
        out_base = torch.where(torch.rand(out_total, generator=gen, device=dev) < spec.x_rate,
                               torch.full_like(out_base, 15), out_base)

    # cut the flanks to the FLANK bases next to the tract (call_locus.py:1144-1146)
    key = out_read * 3 + out_region
    reg_len = torch.bincount(key, minlength=nr * 3)
    reg_start = torch.cumsum(reg_len, 0) - reg_len
    idx_in = torch.arange(out_total, device=dev) - reg_start[key]
    keep2 = torch.where(out_region == 0, (reg_len[key] - 1 - idx_in) < FLANK,
                        torch.where(out_region == 2, idx_in < FLANK, torch.ones_like(idx_in, dtype=torch.bool)))
    out_base = out_base[keep2]
    key = key[keep2]
    lens = torch.bincount(key, minlength=nr * 3).view(nr, 3)
    tot = lens.sum(dim=1)
    seq_off = torch.cumsum(tot, 0) - tot
    est = torch.round(lens[:, 1].double() / mr.double()).long()  # round-half-even, like Python's round()

    ascii_lut = _ASCII.to(dev)
    reads = ascii_lut[out_base]
    motif_flat = ascii_lut[motif[inside]]
    motif_off = torch.cumsum(m, 0) - m
    return dict(reads=reads, motifs=motif_flat, seq_off=seq_off, lens=lens.to(torch.int32), est_cn=est.to(torch.int32),
                read_begin=torch.arange(0, nr + 1, rpl, device=dev, dtype=torch.int64), motif_off=motif_off,
                motif_len=m.to(torch.int32), true_cn=a.to(torch.int32))


# The motif column of the reference's catalogs/pathogenic_assoc.hg38.tsv, rows 3-46 in file order (44 loci; IUPAC motifs
# kept as they are: RAAAT, AAAWK, GCN, CASR, ATTTY, AARRG, TTTYA, AARTA, TRC, TRRAA).  BASELINE config 4 = these + 16 resampled.
PATHOGENIC_MOTIFS = ("RAAAT GCC GGC AAAWK GCC GCN GCA GCA CASR ATTTY CAG AARRG GCC TTTYA GCT TGC GGC GCA GCN AARTA GCCCCG CGG CGG "
                     "CAG GGC TRC GCG GCG GCG GGC ATTTY TRRAA AGC CTG CCG CGT CAG GGGCCT GCGCGGGGCGGG ATTCT GCA GAGAGG GGC GCC").split()
EXPANSION_MOTIFS = PATHOGENIC_MOTIFS  # (name kept for callers of the round-1 generator)
# bases a motif code stands for when a read is synthesised: the reference's own table (strkit/iupac.py:9-21), including its
# quirk that D lists (A, C, T) like H -- a read base G under a D column would score as a mismatch in the reference too
IUPAC_BASES = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "B": "CGT", "D": "ACT", "H": "ACT",
               "V": "ACG", "N": "ACGT"}


def expansion_motifs(n_loci: int, rng) -> list[str]:
    """The 44 catalog motifs in file order, then motifs resampled from them (44 + 16 for 60 loci)."""
    extra = [PATHOGENIC_MOTIFS[int(i)] for i in rng.integers(0, len(PATHOGENIC_MOTIFS), max(0, n_loci - len(PATHOGENIC_MOTIFS)))]
    return (list(PATHOGENIC_MOTIFS) + extra)[:n_loci]


def generate_expansions(n_loci: int = 60, reads_per_locus: int = 50, seed: int = 20261018 + 4000,
                        max_tract: int = 6000, big_lo: int = 200, big_hi: int = 2000, motifs: list[str] | None = None):
    """Config 4: short allele 10-40 copies, expanded allele U{big_lo..big_hi} copies capped at max_tract bases, HiFi
    error channel.  Small set; per-read numpy.  Returns (ReadBatch, list[LocusReads])."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    iupac = IUPAC_BASES
    motifs = expansion_motifs(n_loci, rng) if motifs is None else motifs

    def channel(seq: np.ndarray) -> np.ndarray:
        u = rng.random(seq.shape[0])
        keep = u >= 0.001
        sub = keep & (u < 0.0015)
        seq = np.where(sub, acgt[rng.integers(0, 4, seq.shape[0])], seq)
        ins = rng.random(seq.shape[0]) < 0.001
        counts = keep.astype(np.int64) + ins
        out = np.repeat(seq, counts)
        src = np.repeat(np.arange(seq.shape[0]), counts)
        first = np.concatenate([[True], src[1:] != src[:-1]])
        inserted = ~first | (first & ~keep[src])
        out = np.where(inserted, acgt[rng.integers(0, 4, out.shape[0])], out)
        return np.where(rng.random(out.shape[0]) < 0.001, ord("X"), out).astype(np.uint8)

    loci = []
    for li in range(n_loci):
        motif = motifs[li % len(motifs)]
        m = len(motif)
        small = int(rng.integers(10, 41))
        big = min(int(rng.integers(big_lo, big_hi + 1)), max_tract // m)
        fl = acgt[rng.integers(0, 4, FLANK)]
        fr = acgt[rng.integers(0, 4, FLANK)]
        est, trs, fls, frs = [], [], [], []
        for _ in range(reads_per_locus):
            a = small if rng.random() < 0.5 else big
            unit = np.array([ord(rng.choice(list(iupac[c]))) if c in iupac else ord(c) for c in motif * a],
                            dtype=np.uint8)
            f1, t1, f2 = channel(fl), channel(unit), channel(fr)
            trs.append(t1.tobytes().decode())
            fls.append(f1[-FLANK:].tobytes().decode())
            frs.append(f2[:FLANK].tobytes().decode())
            est.append(int(round(len(t1) / m)))
        loci.append(LocusReads(motif, est, trs, fls, frs))
    return pack_loci(loci), loci
