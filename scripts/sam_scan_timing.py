"""Host side of the SAM -> MIDSV path: the Python packer (read_sam_records + pack_alignments + the cs-level indel fix)
against alignment.scan_sam (hostc/samscan.c) on a synthetic SAM file with SEQ / QUAL columns like minimap2 writes."""
from __future__ import annotations

import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from cstag_bench import tags  # noqa: E402

from dajin2_b200 import alignment  # noqa: E402


def main(n: int = 20000, ref_len: int = 5000) -> None:
    recs = tags(2000, ref_len, 5)
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "sample.sam"
        with open(path, "w") as f:
            f.write(f"@SQ\tSN:allele\tLN:{ref_len}\n@PG\tID:minimap2\n")
            seq, qual = "A" * ref_len, "I" * ref_len
            for k in range(n):
                q, flag, rname, pos, cs = recs[k % len(recs)]
                f.write(f"{q}_{k:07d}\t{flag}\t{rname}\t{pos}\t60\t{ref_len}M\t*\t0\t0\t{seq}\t{qual}\tNM:i:12\tms:i:9000\tcs:Z:{cs}\n")
        size = path.stat().st_size
        t = time.perf_counter()
        sq, records = alignment.read_sam_records(path)
        (a,) = alignment.pack_alignments(sq, [(q, f, r, p, alignment.revert_deletion_insertion_pairs(cs))
                                              for q, f, r, p, cs in records])
        t_py = time.perf_counter() - t
        alignment.scan_sam(path, fix_indel_pairs=True)
        t = time.perf_counter()
        (b,) = alignment.scan_sam(path, fix_indel_pairs=True)
        t_c = time.perf_counter() - t
        assert a.qnames == b.qnames and (a.aln_pos == b.aln_pos).all()
        print(f"{n} records, {size / 1e6:.0f} MB: Python packer {t_py:.2f} s, scan_sam {t_c:.3f} s "
              f"({size / t_c / 1e9:.2f} GB/s on {alignment._threads()} threads, {t_py / t_c:.0f}x)")


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
