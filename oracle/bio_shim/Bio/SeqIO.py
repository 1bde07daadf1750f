"""TEST INFRASTRUCTURE ONLY — minimal FASTA reader with the attributes the reference uses
(``.id``, ``.description``, ``.seq`` as a plain ``str``; treetool/vcftree.py:149-157,166,171)."""


class _Record:
    __slots__ = ("id", "description", "seq")

    def __init__(self, header, seq):
        self.description = header
        self.id = header.split()[0] if header.split() else ""
        self.seq = seq


def parse(handle_or_path, fmt):
    if fmt != "fasta":
        raise ValueError("the Biopython stand-in only reads FASTA")
    close = False
    if isinstance(handle_or_path, (str, bytes)) or hasattr(handle_or_path, "__fspath__"):
        handle = open(handle_or_path)
        close = True
    else:
        handle = handle_or_path
    try:
        header, chunks = None, []
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if header is not None:
                    yield _Record(header, "".join(chunks))
                header, chunks = line[1:], []
            elif header is not None:
                chunks.append(line.strip())
        if header is not None:
            yield _Record(header, "".join(chunks))
    finally:
        if close:
            handle.close()
