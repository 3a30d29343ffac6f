"""GA population fitness evaluation over StyleSpace directions, sharded by individual (BASELINE configs[3]).

The reference has no working GA evaluator: `GA/` is un-importable (GA/__init__.py:5 names a function that does not exist)
and the hooks in DissimilarDomains/training/training_loop.py:392-434 call undefined helpers.  The workload is defined
from its pieces (SURVEY.md section 8(d) cfg 4):

    individual   = one additive StyleSpace offset `offset[1, C_l]` for every conv / ToRGB layer of a generator built with
                   use_domain_modulation=True, domain_modulation_parametrization='additive'
                   (networks.py:140-160, 515-523); initialised 0.1 * randn (the mutation scale of
                   GA/crossover_mutation.py:17-20)
    fitness_i    = mean over a fixed latent batch of D(G.synthesis(ws; offset_i)) logits -- the quantity
                   apply_genetic_algorithm thresholds on (training_loop.py:412-434); G in eval mode, noise_mode='const'
    sharding     = individual i -> rank i % world (individuals are independent); every rank holds a full G/D replica and
                   the shared latent batch; ONE all_gather of the per-rank fitness slices, nothing else crosses ranks.

    generation   = selection on the gathered [P] vector, the reference's crossover / mutation operators
                   (GA/crossover_mutation.py:4-7 gaussian_crossover, :16-19 dynamic_mutation) on the genomes -- host-side, a few
                   KB per individual, drawn on rank 0 -- and ONE broadcast of the new [P, genome] population (1.9 MB at
                   64 x 7424), so that every rank starts the next evaluation from the same genomes: `next_generation`,
                   `evolve`.
"""
import torch
import torch.distributed as dist


def offset_layers(G):
    """[(name, module)] of every synthesis layer that carries a StyleSpace `offset` parameter, in network order."""
    return [(name, m) for name, m in G.synthesis.named_modules() if isinstance(getattr(m, 'offset', None), torch.nn.Parameter)]


def genome_size(G):
    return sum(m.offset.numel() for _, m in offset_layers(G))


def init_population(G, size, scale=0.1, seed=0):
    """[size, genome] fp32 on the CPU; identical on every rank for a given seed."""
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(size, genome_size(G), generator=gen) * scale


@torch.no_grad()
def load_individual(G, genome):
    """Write one genome [genome_size] into the per-layer offsets of G."""
    pos = 0
    for _, m in offset_layers(G):
        n = m.offset.numel()
        m.offset.copy_(genome[pos:pos + n].reshape(m.offset.shape).to(m.offset.device))
        pos += n
    assert pos == genome.numel(), 'genome length does not match the generator'


def shard_indices(population_size, rank, world):
    """Individuals of this rank: i % world == rank."""
    return list(range(rank, population_size, world))


def gather_fitness(local, population_size, rank, world, device):
    """Per-rank fitness slices -> the full [P] vector on every rank: ONE all_gather (padded to equal length)."""
    per = (population_size + world - 1) // world
    buf = torch.full([per], float('nan'), device=device)
    if len(local):
        buf[:len(local)] = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=device).reshape([]) for v in local])
    if world == 1:
        parts = [buf]
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    out = torch.empty([population_size], device=device)
    for r in range(world):
        idx = shard_indices(population_size, r, world)
        out[idx] = parts[r][:len(idx)]
    return out


@torch.no_grad()
def default_fitness(G, D, ws, c):
    """mean_b D(G.synthesis(ws_b; current offsets))."""
    img = G.synthesis(ws, noise_mode='const')
    return D(img, c).mean()


_graph_cache = dict()       # (id(G), id(D), fitness_fn, ws shape, device) -> (graph, static_ws, static_c, static_out)


@torch.no_grad()
def evaluate_population(G, D, population, z, c=None, rank=0, world=1, fitness_fn=None, cuda_graph=True):
    """Fitness [P] of every individual, identical on all ranks.

    G, D: this rank's replicas (G built with additive domain modulation); population: [P, genome] (same on every rank);
    z: the shared latent batch [B, z_dim] on this rank's device.  `fitness_fn(G, D, ws, c)` defaults to the mean logit.

    One individual is ~130 library launches plus a few hundred small torch ops for ~10 ms of device work at 256^2, partly
    launch-bound from Python.  With `cuda_graph=True` (CUDA only) the evaluation of ONE individual is captured into a CUDA graph
    after an eager warm-up and replayed for every other individual -- only the offsets (device tensors, updated in place between
    replays) change -- and the captured graph is kept for the next call with the same networks and latent shape (the next
    generation), so its ~0.1 s capture is paid once.  The same kernels run either way and the results are identical (measured:
    82 -> 97 individuals/s at 256^2 paper256, batch 8, one B200; `bench.py --workload ga`).
    """
    fitness_fn = fitness_fn or default_fitness
    device = z.device
    if c is None:
        c = torch.zeros([z.shape[0], 0], device=device)
    was_training = (G.training, D.training)
    G.eval(); D.eval()
    ws = G.mapping(z, c)                                  # shared by all individuals: offsets live in the synthesis layers only
    mine = shard_indices(population.shape[0], rank, world)
    local = []
    graph = static_out = None
    if cuda_graph and device.type == 'cuda' and len(mine) >= 2:
        key = (id(G), id(D), fitness_fn, tuple(ws.shape), tuple(c.shape), str(device))
        entry = _graph_cache.get(key)
        if entry is None:
            static_ws, static_c = ws.clone(), c.clone()
            load_individual(G, population[mine[0]])
            side = torch.cuda.Stream(device)              # warm-up on a side stream, as graph capture requires
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                fitness_fn(G, D, static_ws, static_c)     # plugin / attribute / allocator initialisation happens here, not in the capture
            torch.cuda.current_stream(device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fitness_fn(G, D, static_ws, static_c)
            entry = _graph_cache[key] = (graph, static_ws, static_c, static_out)
        graph, static_ws, static_c, static_out = entry
        static_ws.copy_(ws); static_c.copy_(c)
    for i in mine:
        load_individual(G, population[i])
        if graph is not None:
            graph.replay()
            local.append(static_out.clone())
        else:
            local.append(fitness_fn(G, D, ws, c))
    out = gather_fitness(local, population.shape[0], rank, world, device)
    G.train(was_training[0]); D.train(was_training[1])
    return out


# ----------------------------------------------------------------------------
# One GA generation on top of the evaluation: selection, the reference's operators, one broadcast.

def gaussian_crossover(parent1, parent2, generator=None):
    """GA/crossover_mutation.py:4-7: child = mu * p1 + (1 - mu) * p2 with mu ~ N(0, 1) per gene."""
    mu = torch.randn(parent1.shape, generator=generator, dtype=parent1.dtype)
    return mu * parent1 + (1 - mu) * parent2


def dynamic_mutation(genomes, mutation_rate=0.1, generator=None):
    """GA/crossover_mutation.py:16-19: genomes + mutation_rate * N(0, 1)."""
    return genomes + mutation_rate * torch.randn(genomes.shape, generator=generator, dtype=genomes.dtype)


def next_generation(population, fitness, rank=0, world=1, elite=None, mutation_rate=0.1, seed=0, device=None):
    """[P, genome] -> [P, genome]: the `elite` fittest individuals survive unchanged (default P // 4, at least one), every other
    slot is dynamic_mutation(gaussian_crossover(a, b)) of two distinct parents drawn uniformly from the fitter half.  The random
    draws happen on rank 0 (seeded: a run is reproducible whatever the number of ranks) and the result is broadcast -- the one
    collective of a generation besides the all_gather of the fitness.  `device`: where the broadcast buffer lives (the rank's GPU
    under NCCL; None = CPU / gloo).  NaN fitness (an individual that diverged) ranks last."""
    P = int(population.shape[0])
    assert fitness.shape == (P,)
    elite = max(1, P // 4) if elite is None else int(elite)
    assert 1 <= elite <= P
    new = torch.empty_like(population)
    if rank == 0:
        gen = torch.Generator().manual_seed(seed)
        order = torch.argsort(torch.nan_to_num(fitness.detach().float().cpu(), nan=float('-inf')), descending=True)
        parents = order[:max(2, P // 2)] if P >= 2 else order
        new[:elite] = population[order[:elite]]
        n_child = P - elite
        if n_child:
            ia = torch.randint(len(parents), [n_child], generator=gen)
            if len(parents) > 1:                                 # a distinct second parent: a non-zero cyclic shift
                ib = (ia + torch.randint(1, len(parents), [n_child], generator=gen)) % len(parents)
            else:
                ib = ia
            children = gaussian_crossover(population[parents[ia]], population[parents[ib]], generator=gen)
            new[elite:] = dynamic_mutation(children, mutation_rate, generator=gen)
    if world > 1:
        buf = new.to(device) if device is not None else new
        dist.broadcast(buf, src=0)
        new = buf.to(population.device)
    return new


def evolve(G, D, population, z, generations, rank=0, world=1, seed=0, **kw):
    """`generations` rounds of evaluate_population -> next_generation; returns (final population, [fitness per generation])."""
    history = []
    for g in range(int(generations)):
        fit = evaluate_population(G, D, population, z, rank=rank, world=world, **{k: v for k, v in kw.items() if k in ('c', 'fitness_fn', 'cuda_graph')})
        history.append(fit.detach().cpu())
        population = next_generation(population, fit, rank=rank, world=world, seed=seed + g, device=(z.device if z.device.type == 'cuda' else None),
                                     **{k: v for k, v in kw.items() if k in ('elite', 'mutation_rate')})
    return population, history
