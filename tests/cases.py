"""Option sets of the golden fixtures (kept in sync with tests/golden/make_golden.py)."""
CASES = {
    "default": ([], "g1_reads"),
    "m200": (["-m", "200"], "g1_reads"),
    "N_n2": (["-N", "-n", "2"], "g1_reads"),
    "R2": (["-R", "2"], "g1_reads"),
    "o3": (["-o", "3"], "g1_reads"),
    "q20": (["-q", "20"], "g1_reads"),
    "L_e3": (["-L", "-o", "2", "-e", "3"], "g1_reads"),
    "stress": (["-n", "4", "-o", "2", "-e", "10", "-l", "32", "-k", "2"], "g1_reads"),
    "n001": (["-n", "0.01"], "g1_reads"),
    "MOE": (["-M", "2", "-O", "5", "-E", "2"], "g1_reads"),
    "i0_d3": (["-i", "0", "-d", "3", "-o", "2"], "g1_reads"),
    "l20_k1": (["-l", "20", "-k", "1"], "g1_reads"),
    "c": (["-c"], "g1_reads"),
    "short_o3": (["-o", "3"], "g1_short"),
    "short_default": ([], "g1_short"),
}
