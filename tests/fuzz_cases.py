"""Random small alignments for differential testing (test infrastructure).

Each case is SAM text over one or two short contigs with reads whose CIGARs mix every op the BAM
alphabet has (M I D N S H P = X), legal and illegal placements (POS 0, overhanging either contig end,
SEQ shorter/longer than the CIGAR's query length), bases in upper/lower case including N and, sometimes,
IUPAC codes -- i.e. everything SURVEY.md Appendix A lists, in random combination."""
import random

OPS = "MIDNSHP=X"


def random_case(seed):
    rng = random.Random(seed)
    n_contigs = rng.choice([1, 1, 1, 2])
    lens = [rng.choice([1, 2, 7, 20, 33, 64, 150]) for _ in range(n_contigs)]
    lines = ["@HD\tVN:1.6"] + ["@SQ\tSN:c%d\tLN:%d" % (i, L) for i, L in enumerate(lens)]
    exotic = rng.random() < 0.25
    n_reads = rng.randint(1, 30)
    for k in range(n_reads):
        c = rng.randrange(n_contigs)
        L = lens[c]
        n_ops = rng.choice([1, 1, 1, 2, 3, 4, 6])
        ops = []
        for j in range(n_ops):
            op = rng.choice("MMMMMMIDSSNHP=X") if n_ops > 1 else rng.choice("MMMMMM=XSI")
            ops.append((rng.randint(0 if rng.random() < 0.05 else 1, 9), op))
        qlen = sum(n for n, op in ops if op in "MIS=X")
        style = rng.random()
        if style < 0.08:
            slen = max(0, qlen - rng.randint(1, 3))       # SEQ shorter than the CIGAR says
        elif style < 0.14:
            slen = qlen + rng.randint(1, 3)               # longer
        else:
            slen = qlen
        alphabet = "ACGTacgtNn" + ("RYKM=" if exotic and rng.random() < 0.5 else "")
        seq = "".join(rng.choice(alphabet) for _ in range(slen)) or "*"
        where = rng.random()
        if where < 0.06:
            pos = 0
        elif where < 0.16:
            pos = max(1, L - rng.randint(0, 4))           # near / over the end
        elif where < 0.2:
            pos = L + rng.randint(1, 3)                   # beyond the contig
        else:
            pos = rng.randint(1, max(1, L))
        flag = rng.choice([0, 0, 0, 16, 256, 2048, 1024, 4])
        cigar = "".join("%d%s" % (n, op) for n, op in ops) if rng.random() > 0.03 else "*"
        rname = "c%d" % c if rng.random() > 0.03 else "*"
        lines.append("r%d\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t*" % (k, flag, rname, pos, cigar, seq))
    return "\n".join(lines) + "\n"
