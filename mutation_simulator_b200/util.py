"""Host utilities with the reference's names and messages (util.py:17-91)."""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

from .colors import Colors
from .fasta import Fasta


class FastaDuplicateHeaderError(Exception):
    """Raised when a Fasta contains at least one duplicate header."""


def _paint(text: str, code: str, no_color: bool) -> str:
    return text if no_color else f"{code}{text}{Colors.norm}"


def exit_with_error(e: Exception, no_color: bool):
    print(_paint(f"ERROR: {e}", Colors.error, no_color), file=sys.stderr)
    sys.exit(1)


def format_warning(msg: str, no_color: bool) -> str:
    return _paint(f"WARNING: {msg}", Colors.warn, no_color)


def print_warning(msg: str, no_color: bool):
    print(format_warning(msg, no_color), file=sys.stderr)


def print_success(msg, no_color):
    print(_paint(msg, Colors.ok, no_color))


def get_md5(fname) -> str:
    h = hashlib.md5()
    with open(fname, "rb") as fh:
        while True:
            chunk = fh.read(1 << 20)
            if not chunk:
                break
            h.update(chunk)
    return h.hexdigest()


def load_fasta(fname, device=None) -> Fasta:
    """Loads a FASTA file (util.py:77-91).  With a CUDA device index the file is ingested on that GPU and the returned
    object carries the engine holding the resident genome; without, it is parsed on the host and the engines upload."""
    path = Path(fname)
    try:
        return Fasta(str(path.resolve()), device=device)
    except ValueError:
        raise FastaDuplicateHeaderError(f"Fasta {fname} contains duplicate header")
