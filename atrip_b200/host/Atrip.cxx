// Host side of the B200 build of atrip: atrip::Atrip::init / run<F> over the C-ABI device engine
// (include/atrip_b200.h).  Mirrors what the reference's orchestrator does at its boundary
// (src/atrip/Atrip.cxx:54-63, 65-1133) -- sizes from the epsilon tensors, replicated small
// tensors, slicing of the four big tensors, tuple distribution, checkpoint, progress callback,
// max_iterations, energy reduction and sign -- while the per-tuple work, the slice stores and
// the tuple schedule live on the device.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <memory>
#include <vector>

#include <atrip/Atrip.hpp>
#include <atrip/Checkpoint.hpp>
#include <atrip_b200.h>

namespace atrip {

size_t Atrip::rank = 0;
size_t Atrip::np = 1;
MPI_Comm Atrip::communicator;
std::map<std::string, double> Atrip::chrono;
IterationDescriptor IterationDescription::descriptor;

void register_iteration_descriptor(IterationDescriptor d) { IterationDescription::descriptor = d; }

// reference Atrip.cxx:54-63
void Atrip::init(MPI_Comm world) {
  Atrip::communicator = world;
  int r = 0, n = 1;
  MPI_Comm_rank(world, &r);
  MPI_Comm_size(world, &n);
  Atrip::rank = (size_t)r;
  Atrip::np = (size_t)n;
}

namespace {

struct Seconds {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double operator()() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

void ok(int rc, const char *what) {
  if (rc != 0) throw std::string("atrip_b200: ") + what + ": " + atrip_b200_last_error();
}

// full column-major contents of a CTF tensor on this rank.  The dense single-process shim
// exposes its storage; a real (distributed) CTF tensor is gathered with read_all, which is what
// the reference does for the replicated tensors (Atrip.cxx:179-181).
// For F = Complex `ptr` addresses interleaved (re, im) doubles -- std::complex<double>'s own memory
// layout, which is how the C-ABI takes complex tensors (atrip_b200_config.field = 1).
template <typename F>
struct HostView {
  const double *ptr = nullptr;
  std::vector<F> owned;
};
template <typename F>
HostView<F> view(CTF::Tensor<F> *t) {
  HostView<F> v;
  if (!t) return v;
#ifdef ATRIP_B200_DENSE_CTF_HPP
  v.ptr = reinterpret_cast<const double *>(t->data);
#else
  int64_t n = 1;
  for (int i = 0; i < t->order; i++) n *= t->lens[i];
  v.owned.resize((size_t)n);
  t->read_all(v.owned.data());
  v.ptr = reinterpret_cast<const double *>(v.owned.data());
#endif
  return v;
}

struct EngineHandle {
  atrip_b200_ctx *ctx = nullptr;
  ~EngineHandle() {
    if (ctx) atrip_b200_destroy(ctx);
  }
};

// Several ranks: every rank slices, out of the (distributed) CTF tensor, only the sources it owns and
// hands them to the engine one batch of slices at a time -- what SliceUnion<F>::init does with
// slice_into_buffer / slice_into_vector (SliceUnion.cxx:305-332, Unions.hpp:21-75, 96-112, ...).
// CTF::Tensor::slice is collective over the tensor's world, so every rank makes the same number of
// slice calls: ranks with fewer sources repeat their first one, like the reference's source padding
// (SliceUnion.cxx:322-326).  kind: atrip_b200_upload_slices; box(x, y, low, up) = the source box.
template <typename F, typename Box>
void upload_owned(atrip_b200_ctx *ctx, CTF::Tensor<F> *origin, int kind, std::vector<int> const &slice_lens, Box box) {
  const int np = (int)Atrip::np, me = (int)Atrip::rank;
  // (x, y) lists: mine, and the longest list of any rank (for the padding)
  int64_t n_mine = 0, n_max = 0;
  const int64_t nv = kind == 101 || kind == 111 ? origin->lens[3] : origin->lens[0];
  std::vector<int64_t> xy;
  for (int r = 0; r < np; r++) {
    const int64_t n = atrip_b200_host_owned_slices(kind, nv, r, np, nullptr, 0);
    if (n < 0) throw std::string("atrip_b200: owned_slices: ") + atrip_b200_last_error();
    n_max = std::max(n_max, n);
    if (r == me) n_mine = n;
  }
  xy.resize((size_t)std::max<int64_t>(1, 2 * n_mine));
  atrip_b200_host_owned_slices(kind, nv, me, np, xy.data(), n_mine);
  CTF::World self(MPI_COMM_SELF);
  std::vector<int> syms(slice_lens.size(), NS), zero(slice_lens.size(), 0);
  CTF::Tensor<F> to((int)slice_lens.size(), slice_lens.data(), syms.data(), self);
  size_t per = 1;
  for (int l : slice_lens) per *= (size_t)l;
  // batches of up to ~32 MB of slices per engine call
  const size_t per_call = std::max<size_t>(1, (size_t)(32u << 20) / (per * sizeof(F)));
  std::vector<F> buf(per * std::min<size_t>(per_call, (size_t)std::max<int64_t>(1, n_mine)));
  std::vector<int64_t> bxy;
  size_t filled = 0;
  auto flush = [&] {
    if (!filled) return;
    ok(atrip_b200_upload_slices(ctx, kind, (int64_t)filled, bxy.data(), reinterpret_cast<const double *>(buf.data())),
       "upload_slices");
    filled = 0;
    bxy.clear();
  };
  for (int64_t it = 0; it < n_max; it++) {
    const bool padding = it >= n_mine;
    if (padding && n_mine == 0) {  // nothing owned at all: still take part in the collective slice
      std::vector<int> low(origin->order, 0), up(origin->order, 1);
      std::vector<int> tl(slice_lens.size(), 0), tu(slice_lens.size(), 1);
      to.slice(tl.data(), tu.data(), F(0), *origin, low.data(), up.data(), F(1));
      continue;
    }
    const int64_t x = xy[2 * (padding ? 0 : it)], y = xy[2 * (padding ? 0 : it) + 1];
    std::vector<int> low, up;
    box((int)x, (int)y, low, up);
    to.slice(zero.data(), slice_lens.data(), F(0), *origin, low.data(), up.data(), F(1));
    if (padding) continue;
    std::copy(to.data, to.data + per, buf.begin() + filled * per);
    bxy.push_back(x);
    bxy.push_back(y);
    if (++filled == per_call) flush();
  }
  flush();
}

// Atrip::run<F> for F = double and F = Complex: the engine computes both fields, the host logic is
// the same (the reference's is one template, Atrip.cxx:65-1133)
template <typename F>
Atrip::Output run_on_engine(Atrip::Input<F> const &in) {
  using Output = Atrip::Output;
  constexpr bool is_cplx = traits::is_complex<F>::value;
  if (!in.ei || !in.ea || !in.Tph || !in.Tpphh || !in.Vpphh || !in.Vhhhp || !in.Vppph)
    throw std::string("atrip: epsilon_i, epsilon_a, Tai, Tabij, Vabij, Vijka and Vabci are all required");
  const size_t No = (size_t)in.ei->lens[0], Nv = (size_t)in.ea->lens[0];  // Atrip.cxx:72-73
  LOG(0, "Atrip") << "No: " << No << "\n";
  LOG(0, "Atrip") << "Nv: " << Nv << "\n";
  LOG(0, "Atrip") << "np: " << Atrip::np << "\n";
  Atrip::chrono.clear();
  Seconds total;

  // one rank per GPU, device = rank % cards (Atrip.cxx:84-91, 119)
  const int ncards = atrip_b200_device_count();
  if (ncards == 0)
    throw std::string("atrip: no CUDA device visible to this rank; the B200 build has no CPU path");
  const bool with_J = in.Jppph && in.Jhhhp;  // both needed for the (cT) pass (Atrip.cxx:338-362)
  atrip_b200_config cfg{};
  cfg.device = (int32_t)(Atrip::rank % (size_t)ncards);
  cfg.rank = (int32_t)Atrip::rank;
  cfg.nranks = (int32_t)Atrip::np;
  cfg.with_J = with_J;
  cfg.No = (int64_t)No;
  cfg.Nv = (int64_t)Nv;
  cfg.batch_tuples = 0;
  cfg.field = is_cplx ? 1 : 0;
  // several ranks: every GPU stores the slices it owns and fetches the rest from its peers over
  // NCCL (the reference's SliceUnion sources + MPI fetches); ATRIP_B200_REPLICATE=1 keeps a full
  // replica per GPU instead (small problems)
  const char *rep = std::getenv("ATRIP_B200_REPLICATE");
  cfg.resident = (Atrip::np == 1 || (rep && rep[0] == '1')) ? 1 : 0;
  const char *tr = std::getenv("ATRIP_B200_TRANSPORT");  // 1 = NCCL send/recv, 2 = P2P copy engines
  cfg.transport = tr ? std::atoi(tr) : 0;
  EngineHandle eng;
  ok(atrip_b200_create(&eng.ctx, &cfg), "create");
  LOG(0, "Atrip") << "engine: " << atrip_b200_version() << " on device " << cfg.device << "\n";
  if (Atrip::np > 1) {  // one NCCL communicator over the ranks of Atrip::init's MPI communicator
    unsigned char id[128] = {0};
    if (Atrip::rank == 0) ok(atrip_b200_comm_unique_id(id), "comm_unique_id");
    MPI_Bcast(id, 128, MPI_BYTE, 0, Atrip::communicator);
    ok(atrip_b200_comm_init(eng.ctx, id), "comm_init");
  }

  {  // replicated tensors (Atrip.cxx:176-215); Tai is negated for the ijkabc algorithm (:183-187)
    Seconds t;
    auto ei = view(in.ei), ea = view(in.ea), tph = view(in.Tph);
    ok(atrip_b200_set_epsilon(eng.ctx, ei.ptr, ea.ptr), "set_epsilon");
    if (in.ijkabc) {
      std::vector<double> neg(tph.ptr, tph.ptr + (is_cplx ? 2 : 1) * No * Nv);
      for (auto &x : neg) x = -x;
      ok(atrip_b200_set_Tai(eng.ctx, neg.data()), "set_Tai");
    } else {
      ok(atrip_b200_set_Tai(eng.ctx, tph.ptr), "set_Tai");
    }
    // the four big tensors -> HBM stores (replaces the five SliceUnion ctors, Atrip.cxx:277-332)
    const int iNo = (int)No, iNv = (int)Nv;
    if (Atrip::np == 1) {
      // one rank owns everything: stream the whole tensors through the engine's bulk ingest
      {
        auto v = view(in.Vppph);
        ok(atrip_b200_load_Vabci(eng.ctx, v.ptr), "load_Vabci");
      }
      if (in.delete_Vppph) delete in.Vppph;  // Atrip.cxx:310
      {
        auto v = view(in.Tpphh);
        ok(atrip_b200_load_Tabij(eng.ctx, v.ptr), "load_Tabij");
      }
      {
        auto v = view(in.Vpphh);
        ok(atrip_b200_load_Vabij(eng.ctx, v.ptr), "load_Vabij");
      }
      {
        auto v = view(in.Vhhhp);
        ok(atrip_b200_load_Vijka(eng.ctx, v.ptr), "load_Vijka");
      }
      if (with_J) {
        auto j1 = view(in.Jhhhp), j2 = view(in.Jppph);
        ok(atrip_b200_load_Jijka(eng.ctx, j1.ptr), "load_Jijka");
        ok(atrip_b200_load_Jabci(eng.ctx, j2.ptr), "load_Jabci");
      }
    } else {
      // several ranks: each slices and uploads only the sources it owns (SliceUnion.cxx:305-332); the
      // boxes are the reference's (Unions.hpp:96-112 TAPHH, 134-151 HHHA, 177-196 ABPH, 219-236 ABHH,
      // 258-277 TABHH)
      auto ab_box = [&](int d2, int d3) {
        return [=](int x, int y, std::vector<int> &low, std::vector<int> &up) {
          low = {x, y, 0, 0};
          up = {x + 1, y + 1, d2, d3};
        };
      };
      auto hhha_box = [&](int x, int, std::vector<int> &low, std::vector<int> &up) {
        low = {0, 0, 0, x};
        up = {iNo, iNo, iNo, x + 1};
      };
      upload_owned<F>(eng.ctx, in.Vppph, 200, {iNv, iNo}, ab_box(iNv, iNo));  // ABPH
      if (with_J) upload_owned<F>(eng.ctx, in.Jppph, 210, {iNv, iNo}, ab_box(iNv, iNo));
      if (in.delete_Vppph) delete in.Vppph;  // Atrip.cxx:310
      upload_owned<F>(eng.ctx, in.Tpphh, 100, {iNv, iNo, iNo},
                      [&](int x, int, std::vector<int> &low, std::vector<int> &up) {  // TAPHH
                        low = {x, 0, 0, 0};
                        up = {x + 1, iNv, iNo, iNo};
                      });
      upload_owned<F>(eng.ctx, in.Tpphh, 201, {iNo, iNo}, ab_box(iNo, iNo));  // TABHH
      upload_owned<F>(eng.ctx, in.Vpphh, 202, {iNo, iNo}, ab_box(iNo, iNo));  // ABHH
      upload_owned<F>(eng.ctx, in.Vhhhp, 101, {iNo, iNo, iNo}, hhha_box);     // HHHA
      if (with_J) upload_owned<F>(eng.ctx, in.Jhhhp, 111, {iNo, iNo, iNo}, hhha_box);
    }
    Atrip::chrono["slicing"] = t();
  }

  {  // tuple distribution (Atrip.cxx:383-399)
    Seconds t;
    ok(atrip_b200_build_tuples(eng.ctx, in.tuples_distribution == Atrip::Input<F>::GROUP_AND_SORT ? 1 : 0),
       "build_tuples");
    Atrip::chrono["tuples:build"] = t();
  }
  const size_t n_iterations = (size_t)atrip_b200_num_tuples(eng.ctx);
  LOG(0, "Atrip") << "#iterations: " << n_iterations << "\n";
  const double doubles_flops = atrip_b200_flops_per_tuple(eng.ctx) / 1e9;  // GF per tuple, Atrip.cxx:578-580

  if (in.rank_round_robin)
    LOG(0, "Atrip") << "note: rank_round_robin has no effect, every GPU is its own node in the slice ownership map\n";
  if (in.blocking)
    LOG(0, "Atrip") << "note: blocking has no effect, slices are fetched one batch ahead on side streams\n";

  // sum over the ranks: ncclAllReduce inside the engine (replaces MPI_Reduce, Atrip.cxx:1094-1107)
  auto sum_ranks = [&](double *v, int n) {
    if (Atrip::np > 1) ok(atrip_b200_allreduce(eng.ctx, v, n), "allreduce");
  };

  // checkpoint (Atrip.cxx:586-621): resume at the stored iteration, rank 0 seeds its energy.  The file
  // stores ranks-per-node and nodes (Atrip.cxx:716-722); here every GPU is a node of its own.
  Output local{0, 0};
  size_t first_iteration = 0;
  const size_t checkpoint_mod = in.checkpoint_at_every_iteration != 0
                                    ? in.checkpoint_at_every_iteration
                                    : (size_t)(n_iterations * in.checkpoint_at_percentage / 100);
  bool resumed = false;
  if (in.read_checkpoint_if_exists) {
    std::ifstream fin(in.checkpoint_path);
    if (fin.is_open()) {
      LOG(0, "Atrip") << "Reading checkpoint from " << in.checkpoint_path << "\n";
      const Checkpoint c = read_checkpoint(fin);
      // the iteration counts tuples of THIS distribution over THIS many ranks: anything else is
      // another calculation's file (the reference leaves these checks as TODOs, Atrip.cxx:599-610)
      if (c.no != No || c.nv != Nv || c.iteration > n_iterations || c.nranks * c.nnodes != Atrip::np)
        throw std::string("atrip: checkpoint ") + in.checkpoint_path +
            " does not belong to this calculation (No, Nv, number of ranks or iteration differ)";
      first_iteration = c.iteration;
      if (Atrip::rank == 0) {  // stored energies are the physical ones
        local.energy = -c.energy;
        local.ct_energy = c.has_ct ? -c.ct_energy : -c.energy;
      }
      resumed = true;
      LOG(0, "Atrip") << "iteration from checkpoint " << first_iteration << "\n";
    }
  }

  // iteration range: the reference leaves its loop after iteration index max_iterations
  // (Atrip.cxx:1052-1056), i.e. max_iterations + 1 tuples are processed
  size_t last = n_iterations;
  if (in.max_iterations != 0) last = std::min(n_iterations, in.max_iterations + 1);

  // reports every iteration_mod iterations or percentage_mod percent (Atrip.cxx:402-405, 734)
  size_t report_mod = 0;
  if (in.percentage_mod > 0) report_mod = std::max<size_t>(1, n_iterations * (size_t)in.percentage_mod / 100);
  else if (in.iteration_mod > 0) report_mod = (size_t)in.iteration_mod;
  const size_t ckpt_mod = in.writeCheckpoint ? checkpoint_mod : 0;
  // a device call runs up to the next report or checkpoint boundary (multiples of the two moduli)
  auto next_boundary = [&](size_t it) {
    size_t b = last;
    if (report_mod) b = std::min(b, (it / report_mod + 1) * report_mod);
    if (ckpt_mod) b = std::min(b, (it / ckpt_mod + 1) * ckpt_mod);
    return b;
  };

  Seconds loop;
  double device_ms = 0, tuples_done = 0;
  bool wrote_checkpoint = false;
  for (size_t it = first_iteration; it < last;) {
    const size_t n = next_boundary(it) - it;
    double e = 0, ect = 0;
    ok(atrip_b200_run(eng.ctx, (int64_t)it, (int64_t)n, &e, &ect), "run");
    double tm[6];
    atrip_b200_last_timing(eng.ctx, tm);
    device_ms += tm[0];
    tuples_done += tm[5];
    local.energy += e;
    local.ct_energy += ect;
    it += n;
    if (in.barrier) MPI_Barrier(Atrip::communicator);
    if (report_mod && (it % report_mod == 0 || it == last)) {
      if (IterationDescription::descriptor) IterationDescription::descriptor({it, n_iterations, loop()});
      LOG(0, "Atrip") << "iteration " << it << " [" << 100 * it / std::max<size_t>(1, n_iterations) << "%] ("
                      << (device_ms > 0 ? doubles_flops * tuples_done / (device_ms * 1e-3) : -1) << "GF)\n";
    }
    if (ckpt_mod && it < last && it % ckpt_mod == 0) {
      double g[2] = {local.energy, local.ct_energy};
      sum_ranks(g, 2);
      if (Atrip::rank == 0) {
        Checkpoint c{No, Nv, 1, Atrip::np, -g[0], it, in.rank_round_robin};
        c.ct_energy = -g[1];
        c.has_ct = with_J;
        write_checkpoint(c, in.checkpoint_path);
      }
      wrote_checkpoint = true;
    }
  }
  Atrip::chrono["iterations"] = loop();
  Atrip::chrono["device"] = device_ms * 1e-3;

  // energy reduction and sign (Atrip.cxx:1094-1111)
  double g[2] = {local.energy, local.ct_energy};
  sum_ranks(g, 2);
  Output global{g[0], g[1]};
  if (!in.ijkabc) {
    global.energy = -global.energy;
    global.ct_energy = -global.ct_energy;
  }
  // a run that walked the whole list is finished: its checkpoint must not survive, or the next run
  // in this directory would resume from it and count most tuples twice
  if (last == n_iterations && (wrote_checkpoint || resumed) && in.writeCheckpoint && Atrip::rank == 0)
    std::remove(in.checkpoint_path.c_str());
  Atrip::chrono["total"] = total();

  // the reference leaves std::cout at 15 digits for its caller's "Energy:" line
  LOG(0, "Atrip") << "Energy: " << std::setprecision(15) << std::setw(23) << global.energy << std::endl;
  if (in.chrono)
    for (auto const &p : Atrip::chrono) LOG(1, " ") << p.first << " :: " << p.second << std::endl;
  LOG(0, "atrip:flops(doubles)") << (device_ms > 0 ? tuples_done * doubles_flops / (device_ms * 1e-3) : 0) << "\n";
  LOG(0, "atrip:flops(iterations)") << tuples_done * doubles_flops / std::max(1e-9, Atrip::chrono["iterations"]) << "\n";
  return global;
}

}  // namespace

// the reference's two instantiations (Atrip.cxx:1135-1136)
template <>
Atrip::Output Atrip::run<double>(Atrip::Input<double> const &in) {
  return run_on_engine<double>(in);
}
template <>
Atrip::Output Atrip::run<Complex>(Atrip::Input<Complex> const &in) {
  return run_on_engine<Complex>(in);
}

}  // namespace atrip
