"""`python -m kmerpapa_b200` — same entry point as the reference's `python -m kmerpapa`."""
import sys

from .cli import main

if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
