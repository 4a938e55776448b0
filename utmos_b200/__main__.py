"""Command dispatcher of the B200 build: ``python -m utmos_b200 {convert,select,version} [options]``.

Behaviour follows the reference CLI (utmos/__main__.py:17-47): the first word picks the command, everything after
it is handed to that command untouched; without any argument the overview goes to stderr and the exit status is 0.
The sub-commands are imported on demand, so ``version`` and the overview work without loading the CUDA library.
"""
import importlib
import sys

from utmos_b200 import __version__

# command -> (module, function, one line for the overview)
COMMANDS = {
    "convert": ("utmos_b200.convert", "cvt_main", "VCF -> packed genotype presence (.jl)"),
    "select": ("utmos_b200.select", "select_main", "greedy maximum-coverage sample selection on the GPU"),
    "version": (None, None, "print the version"),
}


def overview():
    lines = [f"Utmos v{__version__} (B200 build): pick the samples that together carry the most variants", "",
             "usage: utmos CMD [OPTIONS ...]", ""]
    width = max(len(name) for name in COMMANDS)
    lines += [f"    {name.ljust(width)}   {about}" for name, (_, _, about) in COMMANDS.items()]
    lines += ["", "`utmos CMD -h` lists the options of a command."]
    return "\n".join(lines) + "\n"


def run(argv):
    """argv without the program name.  Returns the process exit status."""
    if not argv:
        sys.stderr.write(overview())
        return 0
    name, rest = argv[0], argv[1:]
    if name in ("-h", "--help"):
        sys.stdout.write(overview())
        return 0
    if name not in COMMANDS:
        sys.stderr.write(overview())
        sys.stderr.write(f"\nutmos: error: unknown command {name!r} (choose from {', '.join(COMMANDS)})\n")
        return 2
    if name == "version":
        print(f"Utmos v{__version__}")
        return 0
    module, function, _ = COMMANDS[name]
    getattr(importlib.import_module(module), function)(rest)
    return 0


def main():
    sys.exit(run(sys.argv[1:]))


if __name__ == "__main__":
    main()
