"""Random clip-heavy alignments for differential testing of the --realign path (test infrastructure).

A case is SAM text over one or two contigs in which a "sample" differs from the reference by a few
replaced segments: reads that cross a breakpoint are soft-clipped there (right clips `kM jS` at the left
breakpoint, left clips `jS kM` at the right one), with the clipped bases taken from the replacement
sequence, so that clip-dominant regions (reference kindel/kindel.py:156-275) appear, extend, decay, overlap
in -> / <- pairs and merge by LCS (kindel.py:278-366).  Mixed in: substitutions, Ns, small indels, clips
near the contig ends (inside the masked ends), and breakpoints close to each other."""
import random


def _mutate(rng, s, p):
    return "".join(rng.choice("ACGTN") if rng.random() < p else ch for ch in s)


def clip_case(seed):
    rng = random.Random(10_000 + seed)
    n_contigs = rng.choice([1, 1, 2])
    lines = ["@HD\tVN:1.6\tSO:unsorted"]
    contigs = []
    for c in range(n_contigs):
        L = rng.choice([60, 90, 140, 220, 320])
        ref = "".join(rng.choice("ACGT") for _ in range(L))
        contigs.append((L, ref))
        lines.append("@SQ\tSN:k%d\tLN:%d" % (c, L))
    k = 0
    for c, (L, ref) in enumerate(contigs):
        n_break = rng.choice([0, 1, 1, 2, 3])
        segs = []  # (b1, b2, replacement)
        for _ in range(n_break):
            b1 = rng.randint(1, L - 2)
            b2 = min(L - 1, b1 + rng.choice([0, 1, 3, 8, 15, 30, 60]))
            z = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 4, 9, 17, 30, 45])))
            segs.append((b1, b2, z))
        depth = rng.choice([3, 8, 20, 40])
        n_reads = max(4, depth * L // 30)
        for _ in range(n_reads):
            rl = rng.randint(12, 48)
            start = rng.randint(0, max(0, L - rl))
            end = min(L, start + rl)
            cigar, seq, pos = None, None, start + 1
            for (b1, b2, z) in segs:
                r = rng.random()
                if start < b1 < end and r < 0.75:  # right clip at b1: matches [start, b1), clipped tail from z + ref[b2:]
                    tail = (z + ref[b2:])[: rng.randint(1, 30)]
                    if tail:
                        cigar = "%dM%dS" % (b1 - start, len(tail))
                        seq = ref[start:b1] + tail
                    break
                if start < b2 < end and r < 0.75:  # left clip at b2: clipped head from ref[:b1] + z, matches [b2, end)
                    head = (ref[:b1] + z)[-rng.randint(1, 30):]
                    if head:
                        cigar = "%dS%dM" % (len(head), end - b2)
                        seq = head + ref[b2:end]
                        pos = b2 + 1
                    break
            if cigar is None:
                body = ref[start:end]
                style = rng.random()
                if style < 0.08 and len(body) > 6:   # small deletion
                    cut = rng.randint(2, len(body) - 3)
                    dl = rng.randint(1, 3)
                    if start + cut + dl < end:
                        cigar = "%dM%dD%dM" % (cut, dl, end - start - cut - dl)
                        seq = body[:cut] + body[cut + dl:]
                elif style < 0.16 and len(body) > 6:  # small insertion
                    cut = rng.randint(2, len(body) - 3)
                    ins = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 4)))
                    cigar = "%dM%dI%dM" % (cut, len(ins), len(body) - cut)
                    seq = body[:cut] + ins + body[cut:]
                elif style < 0.22:                   # clips at the contig ends (masked region)
                    cl = rng.randint(1, 12)
                    junk = "".join(rng.choice("ACGT") for _ in range(cl))
                    if rng.random() < 0.5:
                        cigar, seq = "%dS%dM" % (cl, len(body)), junk + body
                    else:
                        cigar, seq = "%dM%dS" % (len(body), cl), body + junk
                if cigar is None:
                    cigar, seq = "%dM" % len(body), body
            seq = _mutate(rng, seq, rng.choice([0.0, 0.01, 0.05]))
            if rng.random() < 0.1:
                seq = seq.lower()
            flag = rng.choice([0, 16, 0, 16, 2048])
            lines.append("q%d\t%d\tk%d\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (k, flag, c, pos, cigar, seq))
            k += 1
    return "\n".join(lines) + "\n"
