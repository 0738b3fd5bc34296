"""Host side of the reference's multi-threaded mode (option `workers` > 1 with a file output): which worker runs which case numbers
and with which seed -- get_threading_mode/3 and run_fuzzing_loop/7, reference src/erlamsa_main.erl:88-111,254-280. The cases themselves
run on the engine, one batch per worker range (eb200_opts.case_stream_seed / case_stream_first)."""


class AS183(object):
    """OTP `random` as erlamsa_rnd drives it (src/erlamsa_rnd.erl:65,72-83): only what deriving the worker seeds needs"""

    def __init__(self, seed):
        self.a = [abs(int(seed[0])) % 30268 + 1, abs(int(seed[1])) % 30306 + 1, abs(int(seed[2])) % 30322 + 1]

    def uniform(self):
        a = self.a
        a[0], a[1], a[2] = a[0] * 171 % 30269, a[1] * 172 % 30307, a[2] * 170 % 30323
        r = a[0] / 30269 + a[1] / 30307 + a[2] / 30323
        return r - int(r)

    def erand(self, n):
        return int(self.uniform() * n) + 1 if n else 0

    def gen_predictable_seed(self):
        return (self.erand(99999), self.erand(99999), self.erand(99999))


def threading_mode(output, n, workers):
    """get_threading_mode/3: None = single-threaded, else [(A, B, Extra)] per worker: cases A..B, then the one extra case Extra (0 = none)"""
    if n == 1 or workers == 1 or output in ("-", "return", "stdout", "stderr"):
        return None
    div, rem = n // workers, n % workers
    tasks = [[a * div, (a + 1) * div - 1, (a + div * workers) if a <= rem else 0] for a in range(workers)]
    tasks[-1][1] = min(tasks[-1][1], n)
    tasks[0][0] = 1
    return [tuple(t) for t in tasks]


def worker_plan(seed, output, n, workers, same_seed=False):
    """-> None (single-threaded) or a list of batches (worker_seed, first_case, n_cases, stream_first): the cases first_case ..
    first_case + n_cases - 1 take the seeds number first_case - stream_first .. of the stream seeded with worker_seed"""
    tasks = threading_mode(output, n, workers)
    if tasks is None:
        return None
    parent = AS183(seed)                                  # erlamsa_rnd:seed(Seed) again, :265
    plan = []
    for a, b, extra in tasks:
        s = tuple(seed) if same_seed else parent.gen_predictable_seed()
        done = 0
        if a >= 1 and b >= a:                             # FuzzingLoop(.., {A, 0}, B, []): {0, _} and N < I end it at once
            plan.append((s, a, b - a + 1, a))
            done = b - a + 1
        if extra:                                         # FuzzingLoop(.., {R, 0}, R, []): one more case, the next seed of the same stream
            plan.append((s, extra, 1, extra - done))
    return plan
