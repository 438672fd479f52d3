"""Drop-in `scape` package for the infer_pa path: `scape.apa_core` resolves to the B200-native host
mirror so result pickles keep the class path `scape.apa_core.Parameters` (apa_core.py:236) and
`scape infer_pa` / `python -m scape infer_pa` keep working.  The other five sub-commands of the
reference CLI (cli.py:25-30) are outside this repository's scope."""
from .cli import cli


def main():
    cli()
