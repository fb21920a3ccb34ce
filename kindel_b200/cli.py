"""`kindel` command line on the B200 engine (restates the argh CLI of reference kindel/cli.py:9-66).

Sub-commands, flags, defaults and output streams follow the reference: `consensus` prints the
REPORT blocks to stderr and one `>name` / sequence pair per contig to stdout (cli.py:30-33),
`weights` / `features` write TSV to stdout (cli.py:44,50), `version` prints `kindel <version>`;
`variants` (in the reference's README only) is an extension, see kindel.variants.
argh derived the flags from the function signatures (first letter as short option unless two
parameters share it); argparse spells the same set out.  Note the CLI default `--min-overlap 7`
(cli.py:13) differs from the API default 9 (kindel.py:492), as in the reference.
"""
from __future__ import annotations

import argparse
import sys

from . import __version__


def consensus(bam_path, realign=False, min_depth=1, min_overlap=7, clip_decay_threshold=0.1, mask_ends=50,
              trim_ends=False, uppercase=False, gpus=None):
    """Infer consensus sequence(s) from alignment in SAM/BAM format"""
    from . import kindel

    res = kindel.bam_to_consensus(bam_path, realign, min_depth, min_overlap, clip_decay_threshold, mask_ends,
                                  trim_ends, uppercase, devices=gpus)
    print("\n".join(res.refs_reports.values()), file=sys.stderr)
    for record in res.consensuses:
        print(f">{record.name}")
        print(record.sequence)


def weights(bam_path, relative=False, confidence=True, confidence_alpha=0.01, gpus=None):
    """Returns table of per-site nucleotide frequencies and coverage"""
    from . import kindel

    kindel.weights(bam_path, relative, confidence, confidence_alpha, devices=gpus).to_csv(sys.stdout, sep="\t", index=False)


def features(bam_path, gpus=None):
    """Returns table of per-site nucleotide frequencies and coverage including indels"""
    from . import kindel

    kindel.features(bam_path, devices=gpus).to_csv(sys.stdout, sep="\t", index=False)


def variants(bam_path, abs_threshold=1, rel_threshold=0.01, only_variants=False, absolute=False):
    """Output variants exceeding specified absolute and relative frequency thresholds"""
    from . import kindel

    kindel.variants(bam_path, abs_threshold, rel_threshold, only_variants, absolute).to_csv(sys.stdout, sep="\t",
                                                                                            index=False)


def plot(bam_path):
    """Plot sitewise soft clipping frequency across reference and genome"""
    from . import kindel

    return kindel.plotly_clips(bam_path)


def version():
    """Show version"""
    return f"kindel {__version__}"


def _add_gpus(p):
    # extension (not in the reference's CLI): shard the pileup over the GPUs of this node
    p.add_argument("--gpus", type=int, default=None,
                   help="number of GPUs of this node to shard the pileup over (default: $KINDEL_GPUS or 1)")


def build_parser() -> argparse.ArgumentParser:
    fmt = argparse.ArgumentDefaultsHelpFormatter
    parser = argparse.ArgumentParser(prog="kindel", formatter_class=fmt)
    sub = parser.add_subparsers(dest="command")

    p = sub.add_parser("consensus", help=consensus.__doc__, description=consensus.__doc__, formatter_class=fmt)
    p.add_argument("bam_path", help="path to SAM/BAM file")
    p.add_argument("-r", "--realign", action="store_true",
                   help="attempt to reconstruct reference around soft-clip boundaries")
    p.add_argument("--min-depth", type=int, default=1, help="substitute Ns at coverage depths beneath this value")
    p.add_argument("--min-overlap", type=int, default=7, help="match length required to close soft-clipped gaps")
    p.add_argument("-c", "--clip-decay-threshold", type=float, default=0.1,
                   help="read depth fraction at which to cease clip extension")
    p.add_argument("--mask-ends", type=int, default=50,
                   help="ignore clip dominant positions within n positions of termini")
    p.add_argument("-t", "--trim-ends", action="store_true",
                   help="trim ambiguous nucleotides (Ns) from sequence ends")
    p.add_argument("-u", "--uppercase", action="store_true", help="close gaps using uppercase alphabet")
    _add_gpus(p)
    p.set_defaults(func=lambda a: consensus(a.bam_path, a.realign, a.min_depth, a.min_overlap,
                                            a.clip_decay_threshold, a.mask_ends, a.trim_ends, a.uppercase, a.gpus))

    p = sub.add_parser("weights", help=weights.__doc__, description=weights.__doc__, formatter_class=fmt)
    p.add_argument("bam_path", help="path to SAM/BAM file")
    p.add_argument("-r", "--relative", action="store_true", help="output relative nucleotide frequencies")
    p.add_argument("-c", "--confidence", action="store_false", default=True,
                   help="calculate confidence interval for consensus")
    p.add_argument("--confidence-alpha", type=float, default=0.01, help="confidence interval alpha value")
    _add_gpus(p)
    p.set_defaults(func=lambda a: weights(a.bam_path, a.relative, a.confidence, a.confidence_alpha, a.gpus))

    p = sub.add_parser("features", help=features.__doc__, description=features.__doc__, formatter_class=fmt)
    p.add_argument("bam_path", help="path to SAM/BAM file")
    _add_gpus(p)
    p.set_defaults(func=lambda a: features(a.bam_path, a.gpus))

    # `variants` is listed by the reference's README (README.md:106-107) but absent from its code: an extension here
    p = sub.add_parser("variants", help=variants.__doc__, description=variants.__doc__, formatter_class=fmt)
    p.add_argument("bam_path", help="path to SAM/BAM file")
    p.add_argument("-a", "--abs-threshold", type=int, default=1, help="absolute frequency above which to call variants")
    p.add_argument("-r", "--rel-threshold", type=float, default=0.01,
                   help="relative frequency (0.0-1.0) above which to call variants")
    p.add_argument("-o", "--only-variants", action="store_true", help="exclude invariant sites from output")
    p.add_argument("--absolute", action="store_true", help="report absolute variant frequencies")
    p.set_defaults(func=lambda a: variants(a.bam_path, a.abs_threshold, a.rel_threshold, a.only_variants, a.absolute))

    p = sub.add_parser("plot", help=plot.__doc__, description=plot.__doc__, formatter_class=fmt)
    p.add_argument("bam_path", help="path to SAM/BAM file")
    p.set_defaults(func=lambda a: plot(a.bam_path))

    p = sub.add_parser("version", help=version.__doc__, description=version.__doc__)
    p.set_defaults(func=lambda a: version())
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if not getattr(args, "func", None):
        parser.print_usage()
        return 1
    out = args.func(args)
    if out is not None:  # argh prints a command's return value
        print(out)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
