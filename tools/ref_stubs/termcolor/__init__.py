"""Import-only stub (tools/gen_golden.py): the reference imports termcolor.cprint at module load."""
def cprint(*a, **k):
    print(*a)
