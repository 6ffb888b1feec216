"""Drop-in ``VehicleRoutingDataset`` (the reference's problems/EVRP.py:9-280).

Same constructor signature and attributes (``static``, ``dynamic``, ``Elevation``, ``distances``, ``slope``),
same ``__getitem__`` tuple arity rule, and the two environment callbacks

    update_dynamic(dynamic, distances, slope, now_idx, chosen_idx) -> (new_dynamic, reward[B,1])
    update_mask(dynamic, distances, slope, chosen_idx)             -> mask[B,S]

which call the bit-exact CUDA step kernels (csrc/evrp_step.cu) instead of ~1,400 ATen ops per step.
When these bound methods are injected into ``evrp_b200.AttentionModel`` the model recognises them
(``_evrp_b200_native``) and runs the whole rollout in its fused persistent kernel instead.

Node layout is [depot, F stations, N customers] (the layout every consumer of the reference assumes; the
reference's own generator has an off-by-one — SURVEY.md §2.1 D4 — which the documented one-line patch P1
removes; this generator reproduces the patched reference's RNG stream on CPU bit for bit).
"""
import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib
from . import policy_math


def parse_cvrplib(path):
    """Minimal CVRPLIB reader (NODE_COORD_SECTION / DEMAND_SECTION / CAPACITY) for the generalisation test of
    utils/functions.py:213-231.  Returns (coords [n,2] int, demands [n] int, capacity int)."""
    coords, demands, capacity, section = [], [], None, None
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            if not line:
                continue
            head = line.split(":")[0].strip().upper()
            if head == "CAPACITY":
                capacity = int(line.split(":")[1])
            elif head in ("NODE_COORD_SECTION", "DEMAND_SECTION", "DEPOT_SECTION", "EOF"):
                section = head
            elif section == "NODE_COORD_SECTION":
                _, x, y = line.split()[:3]
                coords.append((int(float(x)), int(float(y))))
            elif section == "DEMAND_SECTION":
                demands.append(int(line.split()[1]))
    if capacity is None or not coords or len(coords) != len(demands):
        raise ValueError(f"{path}: not a CVRPLIB instance with coordinates, demands and capacity")
    return coords, demands, capacity


# ------------------------------------------------------------------------------------------------
# test-set files: the reference's pickle format, trainer.py:361-381 (writer) / :469-482 (reader),
# utils/data_utils.py:11-31 -- a list of (static [2,S], dynamic [4,S], distances [S,S], slope [S,S]) nested lists,
# the exact files the classical solvers (BaselineAlgorithms/ACO.py:27-34) consume.
# ------------------------------------------------------------------------------------------------
def check_extension(filename):
    import os
    return filename if os.path.splitext(filename)[1] == ".pkl" else filename + ".pkl"


def save_dataset(dataset, filename):
    """``dataset``: a VehicleRoutingDataset (4-tuple items) or an iterable of (static, dynamic, distances, slope)."""
    import os
    import pickle
    rows = []
    for item in (dataset[i] for i in range(len(dataset))) if hasattr(dataset, "__getitem__") else dataset:
        if len(item) != 4:
            raise ValueError("the test-set file stores 4-tuples (static, dynamic, distances, slope)")
        rows.append(tuple(torch.as_tensor(t).detach().cpu().tolist() for t in item))
    d = os.path.split(filename)[0]
    if d and not os.path.isdir(d):
        os.makedirs(d)
    with open(check_extension(filename), "wb") as f:
        pickle.dump(rows, f, pickle.HIGHEST_PROTOCOL)
    return check_extension(filename)


def load_dataset(filename):
    import pickle
    with open(check_extension(filename), "rb") as f:
        return pickle.load(f)


def EVRPDataset(filename=None, num_samples=256, offset=0, device=None):
    """trainer.py:476-482: list of 4-tuples of tensors (on ``device``), ready for a DataLoader / ``AttentionModel``."""
    import os
    assert os.path.splitext(filename)[1] == ".pkl"
    device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return [tuple(torch.tensor(t, dtype=torch.float32, device=device) for t in row)
            for row in load_dataset(filename)[offset:offset + num_samples]]


def synthetic_instances(B, N, F=5, seed=0, t_limit=10., Start_SOC=80., max_load=4, max_demand=4):
    """B synthetic EVRP instances with the distribution and value grid of problems/EVRP.py:43-92 (layout
    [depot, F stations, N customers]): depot in [25,75)^2, station i in quadrant i % 4, customers on the integer grid
    [0,100]^2, demand in {1..max_demand} * 0.25 / max_load, elevation in {0..100} / 1000, dynamic0 = [1, demand,
    Start_SOC, t_limit].  numpy generator (bulk benchmark input; ``VehicleRoutingDataset`` reproduces the reference's
    torch RNG stream).  -> static [B,2,S], dynamic [B,4,S], elevation [B,S] (float32)."""
    rng = np.random.RandomState(seed)
    S = N + F + 1
    f32 = np.float32
    lo = rng.randint(0, 50, (2, B, F)); hi = rng.randint(51, 101, (2, B, F))
    q = np.arange(F) % 4
    stations = np.stack([np.where(q < 2, lo[0], hi[0]), np.where(q % 2 == 0, lo[1], hi[1])], 1).astype(f32)
    depot = rng.randint(25, 75, (B, 2, 1)).astype(f32)
    cust = rng.randint(0, 101, (B, 2, N)).astype(f32)
    static = np.concatenate([depot, stations, cust], 2)
    demand = rng.randint(1, max_demand + 1, (B, S)).astype(f32) * f32(0.25) / f32(max_load)
    demand[:, :F + 1] = 0
    elevation = rng.randint(0, 101, (B, S)).astype(f32) / f32(1000)
    dynamic = np.stack([np.ones((B, S), f32), demand, np.full((B, S), Start_SOC, f32), np.full((B, S), t_limit, f32)], 1)
    return static, dynamic, elevation


def generate_instances(B, N, F=5, seed=0, t_limit=10., Start_SOC=80., max_load=4, max_demand=4, device=None):
    """On-device instance generation: the same distribution and value grid as ``synthetic_instances`` /
    problems/EVRP.py:43-92, produced by one CUDA kernel from a counter-based generator -- no host loop, no host-to-device
    copy.  -> CUDA tensors static [B,2,S], dynamic [B,4,S], elevation [B,S]; feed them as the 3-tuple input."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    S = N + F + 1
    static = torch.empty(B, 2, S, device=dev)
    dynamic = torch.empty(B, 4, S, device=dev)
    elev = torch.empty(B, S, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evrp_generate_instances(int(seed) & 0xFFFFFFFFFFFFFFFF, B, N, F, float(Start_SOC), float(t_limit),
                                                      int(max_load), int(max_demand), _lib.ptr(static), _lib.ptr(dynamic),
                                                      _lib.ptr(elev), _lib.stream_ptr()), "evrp_generate_instances")
    return static, dynamic, elev


class VehicleRoutingDataset(Dataset):
    _evrp_b200_native = True

    def __init__(self, num_samples, input_size, t_limit, Start_SOC, velocity, max_load, max_demand, charging_num,
                 seed, args, objective=None, device=None):
        super().__init__()
        if seed is None:
            seed = np.random.randint(1234567890)
        np.random.seed(seed)
        torch.manual_seed(seed)
        lib_test = bool(getattr(args, "CVRP_lib_test", False))
        if lib_test:
            num_samples = 1
        self.num_samples, self.input_size = num_samples, input_size
        self.max_load, self.max_demand = max_load, max_demand
        self.Start_SOC, self.t_limit, self.velocity = Start_SOC, t_limit, velocity
        self.charging_num = F = charging_num
        self.objective = objective or ("DM" if str(getattr(args, "obj", "EM")).upper().startswith("D") else "EM")
        self.device = torch.device(device) if device is not None else \
            torch.device("cuda" if torch.cuda.is_available() else "cpu")
        n, N = num_samples, input_size
        S = N + 1 + F
        # RNG draw order of problems/EVRP.py:43-46,68,70,72,77 (CPU generator)
        lo_x = torch.randint(0, 50, (n, 1, 20)); hi_x = torch.randint(51, 101, (n, 1, 20))
        lo_y = torch.randint(0, 50, (n, 1, 20)); hi_y = torch.randint(51, 101, (n, 1, 20))
        q = torch.arange(F) % 4                                       # quadrant of station i
        sx = torch.where((q < 2)[None, :], lo_x[:, 0, :F], hi_x[:, 0, :F])
        sy = torch.where((q % 2 == 0)[None, :], lo_y[:, 0, :F], hi_y[:, 0, :F])
        stations = torch.stack([sx, sy], 1).float()                   # [n,2,F]
        if lib_test:
            coords, dem, cap = parse_cvrplib(args.CVRP_lib_path)
            if len(coords) - 1 != N:
                raise ValueError("CVRPlib instance size does not match input_size")
            xy = torch.tensor(coords, dtype=torch.float32).t()[None]  # [1,2,n]
            locations = torch.cat([xy[:, :, :1], stations, xy[:, :, 1:]], 2)
            demands = torch.cat([torch.zeros(1, 1, 1 + F), (torch.tensor(dem[1:]) / float(cap))[None, None]], 2)
        else:
            depot = torch.randint(25, 75, (n, 2, 1))
            customers = torch.randint(0, 101, (n, 2, N))
            locations = torch.cat((depot.float(), stations, customers.float()), 2)
            demands = torch.randint(1, max_demand + 1, (n, 1, S)) * 0.25
            demands = demands / float(max_load)
            demands[:, :, :1 + F] = 0
        self.static = locations
        self.Elevation = torch.randint(0, 101, (n, S)) / 1000
        ones = torch.ones(n, 1, S)
        self.dynamic = torch.cat((ones, demands.float(), ones * Start_SOC, ones * t_limit), 1).float()
        self.distances = self.slope = None
        if self._precomputed():                                       # problems/EVRP.py:85-92
            chunks = [policy_math.pairwise_matrices(self.static[i:i + 8192].to(self.device),
                                                    self.Elevation[i:i + 8192].to(self.device))
                      for i in range(0, n, 8192)]
            self.distances = torch.cat([c[0] for c in chunks], 0)
            self.slope = torch.cat([c[1] for c in chunks], 0)
            self.Elevation = self.Elevation.to(self.device)

    def _precomputed(self):
        return self.input_size <= 20 or self.num_samples <= 12800

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):                                       # problems/EVRP.py:97-102
        if self._precomputed():
            return (self.static[idx], self.dynamic[idx], self.distances[idx], self.slope[idx])
        return (self.static[idx], self.dynamic[idx], self.Elevation[idx])

    # ------------------------------------------------------------------------------------------
    # x / python-scalar convention of the state recipe: 0 = IEEE division (what ATen evaluates on the CPU,
    # pinned by the golden fixtures), 1 = multiplication by fl(1 / scalar); tests/test_reference_gpu.py records which of
    # the two the reference's CUDA run follows
    div_mode = 0

    def _phys(self):
        return _lib.make_phys(float(self.velocity), float(self.Start_SOC), float(self.t_limit), self.max_load,
                              self.objective, div_mode=int(self.div_mode))

    @staticmethod
    def _cuda(t, dtype):
        if not t.is_cuda:
            raise _lib.EvrpLibraryError("evrp_b200 environment kernels need CUDA tensors (no CPU fallback)")
        return t.to(dtype).contiguous()

    def update_dynamic(self, dynamic, distances, slope, now_idx, chosen_idx):
        """problems/EVRP.py:226-280.  EM: (new_dynamic, energy[B,1]); DM: (new_dynamic, distance[B,1], energy[B,1])
        (DM-EVRP/problems/EVRP.py:280)."""
        L = _lib.lib()
        dyn = self._cuda(dynamic.detach(), torch.float32)
        D, G = self._cuda(distances, torch.float32), self._cuda(slope, torch.float32)
        now, ch = self._cuda(now_idx, torch.int64), self._cuda(chosen_idx, torch.int64)
        B, _, S = dyn.shape
        out = torch.empty_like(dyn)
        e = torch.empty(B, 1, device=dyn.device)
        dist = torch.empty(B, 1, device=dyn.device)
        rc = L.evrp_update_dynamic(self._phys(), _lib.ptr(dyn), _lib.ptr(D), _lib.ptr(G), _lib.ptr(now), _lib.ptr(ch),
                                   B, S, self.charging_num, _lib.ptr(out), _lib.ptr(e), _lib.ptr(dist),
                                   _lib.stream_ptr())
        _lib.check(rc, "evrp_update_dynamic")
        if self.objective == "DM":
            return out, dist, e
        return out, e

    def update_mask(self, dynamic, distances, slope, chosen_idx=None):
        """problems/EVRP.py:105-223 -> mask [B,S] float 0/1 (all-zero once the whole batch is finished)."""
        L = _lib.lib()
        dyn = self._cuda(dynamic.detach(), torch.float32)
        D, G = self._cuda(distances, torch.float32), self._cuda(slope, torch.float32)
        ch = self._cuda(chosen_idx, torch.int64)
        B, _, S = dyn.shape
        mask = torch.empty(B, S, device=dyn.device)
        flag = torch.empty(1, dtype=torch.int32, device=dyn.device)
        rc = L.evrp_update_mask(self._phys(), _lib.ptr(dyn), _lib.ptr(D), _lib.ptr(G), _lib.ptr(ch), B, S,
                                self.charging_num, _lib.ptr(mask), _lib.ptr(flag), _lib.stream_ptr())
        _lib.check(rc, "evrp_update_mask")
        return mask
