"""Minimal FASTA reader with kseq semantics (reference src/kseq.h): name = header up to the
first whitespace, comment = the rest, sequence = concatenated lines; gzip transparent."""
import gzip


def read_fasta(path):
    """-> list of (name, comment, seq bytes)."""
    op = gzip.open if path.endswith(".gz") else open
    recs, name, comment, chunks = [], None, None, []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    recs.append((name, comment, b"".join(chunks)))
                hdr = line[1:].decode(errors="replace")
                parts = hdr.split(None, 1)
                name = parts[0] if parts else ""
                comment = parts[1] if len(parts) > 1 else None
                chunks = []
            elif name is not None:
                chunks.append(line)
    if name is not None:
        recs.append((name, comment, b"".join(chunks)))
    return recs
