"""``python -m infomesh_b200`` — the ``infomesh`` command line."""
from infomesh_b200.cli import cli


def main() -> None:
    cli()


if __name__ == "__main__":
    main()
