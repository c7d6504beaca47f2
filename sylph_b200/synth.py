"""Counter-based synthetic workloads.

Everything is a pure function of (seed, index) evaluated with wrapping int64 arithmetic in torch,
so the same bytes come out on CPU and on CUDA, in any chunking.  No dataset is downloaded.

  genome g, position p  : code = mix(SEED_DB + g, p) & 3              (i.i.d. uniform ACGT)
  every genome with g % 100 == 99 is a ~97 %-identity mutant of genome g-1
  reads                  : community of the first `n_comm` genomes with log-normal abundances
                           (sigma 1.5), uniform start, strand flip p=.5, 0.5 % substitutions,
                           2 % exact duplicates of another read, 0.1 % 'N', 5 % lower-case
"""
import math

import torch

SEED_READS = 0x5EED0001
SEED_DB = 0x5EED0002
_M1 = -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687  # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685  # 0x94D049BB133111EB
_ASCII = (65, 67, 71, 84)  # A C G T


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix(a, b):
    """splitmix64-style mixer of two int64 tensors / scalars -> int64 tensor (wrapping)."""
    x = a * _M1 + b
    x = x ^ _lsr(x, 30)
    x = x * _M2
    x = x ^ _lsr(x, 27)
    x = x * _M3
    x = x ^ _lsr(x, 31)
    return x


def _u01(h):
    """int64 hash -> float64 in [0,1) from the top 53 bits"""
    return _lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def genome_codes(gid, pos):
    """2-bit codes of genome(s) `gid` (int or int64 tensor) at int64 tensor positions `pos`.
    Mutant genomes (gid % 100 == 99) copy gid-1 except at ~3 % of positions."""
    gid_t = gid if torch.is_tensor(gid) else torch.full_like(pos, int(gid))
    is_mut = (gid_t % 100) == 99
    src = torch.where(is_mut, gid_t - 1, gid_t)
    base = mix(src + SEED_DB, pos) & 3
    mh = mix(gid_t + SEED_DB + 0x1000000, pos)
    mutate = is_mut & (_u01(mh) < 0.03)
    alt = (base + 1 + (_lsr(mh, 3) % 3)) & 3
    return torch.where(mutate, alt, base)


def _to_ascii(codes):
    lut = torch.tensor(_ASCII, dtype=torch.uint8, device=codes.device)
    return lut[codes]


def db_chunk(g0, g1, genome_len, device="cpu"):
    """ASCII bases of genomes [g0, g1), one contig each -> (uint8 flat, int64 contig offsets)."""
    n = g1 - g0
    pos = torch.arange(genome_len, dtype=torch.int64, device=device)
    out = torch.empty(n * genome_len, dtype=torch.uint8, device=device)
    for i in range(n):
        out[i * genome_len:(i + 1) * genome_len] = _to_ascii(genome_codes(g0 + i, pos))
    off = torch.arange(n + 1, dtype=torch.int64, device=device) * genome_len
    return out, off


def community_cdf(n_comm, seed=SEED_READS, sigma=1.5):
    h = mix(torch.tensor(seed + 0x77), torch.arange(n_comm, dtype=torch.int64))
    h2 = mix(torch.tensor(seed + 0x78), torch.arange(n_comm, dtype=torch.int64))
    u1 = _u01(h).clamp_min(1e-12)
    u2 = _u01(h2)
    z = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)
    w = torch.exp(sigma * z)
    cdf = torch.cumsum(w / w.sum(), 0)
    cdf[-1] = 1.0
    return cdf


def community_ids(n_comm, n_genomes, seed=SEED_READS):
    """n_comm genome ids spread over [0, n_genomes) (a pure function of the seed; duplicates possible): the
    community of a sample that is profiled against a sharded db, so that survivors come from every shard."""
    return _lsr(mix(torch.tensor(seed + 0x79), torch.arange(n_comm, dtype=torch.int64)), 1) % int(n_genomes)


def reads_chunk(r0, r1, read_len=150, n_comm=64, genome_len=4_000_000, seed=SEED_READS, device="cpu",
                cdf=None, comm=None):
    """ASCII bases of reads [r0, r1) (fixed length) -> uint8 tensor of (r1-r0)*read_len bytes.
    comm: int64 tensor of the community's genome ids (default: genomes 0 .. n_comm-1)."""
    if cdf is None:
        cdf = community_cdf(n_comm, seed)
    cdf = cdf.to(device)
    n = r1 - r0
    j = torch.arange(r0, r1, dtype=torch.int64, device=device)
    # 2 % of reads copy the parameters of another read (exact duplicate of its base version)
    hd = mix(torch.tensor(seed + 1, device=device), j)
    is_dup = (_u01(hd) < 0.02) & (j > 0)
    src = torch.where(is_dup, _lsr(mix(torch.tensor(seed + 2, device=device), j), 1) % j.clamp_min(1), j)
    hg = mix(torch.tensor(seed + 3, device=device), src)
    gid = torch.searchsorted(cdf, _u01(hg)).clamp_max(n_comm - 1)
    if comm is not None:
        gid = comm.to(device)[gid]
    start = _lsr(mix(torch.tensor(seed + 4, device=device), src), 1) % (genome_len - read_len + 1)
    rev = (mix(torch.tensor(seed + 5, device=device), src) & 1) == 1
    o = torch.arange(read_len, dtype=torch.int64, device=device)
    # forward read base o comes from genome position start+o; reverse strand: start+len-1-o, complemented
    gpos = torch.where(rev[:, None], start[:, None] + (read_len - 1) - o[None, :], start[:, None] + o[None, :])
    codes = genome_codes(gid[:, None].expand(n, read_len), gpos)
    codes = torch.where(rev[:, None], 3 - codes, codes)
    he = mix(src[:, None] * 1024 + o[None, :], torch.tensor(seed + 6, device=device))
    ue = _u01(he)
    sub = ue < 0.005
    codes = torch.where(sub, (codes + 1 + (_lsr(he, 5) % 3)) & 3, codes)
    asc = _to_ascii(codes)
    hn = mix(src[:, None] * 1024 + o[None, :], torch.tensor(seed + 7, device=device))
    un = _u01(hn)
    asc = torch.where(un < 0.001, torch.full_like(asc, 78), asc)            # 'N'
    lower = (un >= 0.001) & (un < 0.051)
    asc = torch.where(lower, asc + 32, asc)
    return asc.reshape(-1)


def reads(n_reads, read_len=150, n_comm=64, genome_len=4_000_000, seed=SEED_READS, device="cpu",
          chunk=1 << 19, comm=None):
    """-> (uint8 flat buffer padded to a multiple of 16 bytes, int64 offsets[n_reads+1])"""
    cdf = community_cdf(n_comm, seed)
    total = n_reads * read_len
    buf = torch.zeros((total + 15) // 16 * 16 + 64, dtype=torch.uint8, device=device)
    for r0 in range(0, n_reads, chunk):
        r1 = min(n_reads, r0 + chunk)
        buf[r0 * read_len:r1 * read_len] = reads_chunk(r0, r1, read_len, n_comm, genome_len, seed, device, cdf, comm)
    off = torch.arange(n_reads + 1, dtype=torch.int64, device=device) * read_len
    return buf[:total], off


def sketch_db_range(ctx, g0, g1, genome_len=4_000_000, k=31, c=200, chunk=125, device="cuda"):
    """Sketch synthetic genomes [g0, g1) on the device in batches of `chunk` genomes (bases are generated
    on the device, never cross PCIe: SURVEY §8-d config 5) -> one Genomes handle."""
    parts = []
    for a in range(g0, g1, chunk):
        b = min(g1, a + chunk)
        bases, off = db_chunk(a, b, genome_len, device=device)
        goff = torch.arange(b - a + 1, dtype=torch.int64, device=device)
        parts.append(ctx.sketch_genomes(bases, off, goff, k=k, c=c))
        del bases
    g = ctx.concat_genomes(parts)
    for p in parts:
        p.free()
    return g
