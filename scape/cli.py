"""click group like the reference's cli.py:1-31, carrying the one sub-command this repo implements."""
import click

from .apa_core import infer_pa


@click.group()
def cli():
    """SCAPE-APA infer_pa on NVIDIA B200 (libscape_b200)."""


cli.add_command(infer_pa)
