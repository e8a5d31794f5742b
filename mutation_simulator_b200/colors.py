from .defaults import _TTY


class Colors:
    error, warn, ok, norm = ("\033[0;31m", "\033[0;33m", "\033[0;32m", "\033[0m") if _TTY else ("", "", "", "")
