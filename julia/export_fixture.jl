# Oracle pin kit: runs the UNMODIFIED reference (src/scLENS.jl:649-832) with a seeded RNG and writes
#   * every random quantity it consumed (the "draws bundle" of SURVEY.md 8c), and
#   * every entry of the result Dict (:826-830)
# as raw little-endian arrays plus a manifest, for tests/test_julia_fixture.py (which skips when no fixture exists).
#
#   julia --project=/path/to/scLENS julia/export_fixture.jl /path/to/scLENS OUT_DIR [csv.gz] [seed] [device] [n_perturb]
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia.  It is what turns "parity unpinned" into "pinned" the
# day someone runs it; nothing in the product or in the other tests depends on it.
#
# How the draws are recovered without touching the reference: Julia's task-local RNG is deterministic given its seed, and
# the only RNG consumers inside sclens() are, in this order,
#   1. rand(UInt32(1):UInt32(N), nnz), rand(UInt32(1):UInt32(M), nnz)            (:669)
#   2. random_nz(inp_df, rmix=true): shuffle(nz_val) :275, then one sample(1:N, count, replace=false) per gene in
#      keys(countmap(nz_col)) order :244-247
#   3. 5000 x rand(Normal(0, sqrt(1/nm)), nm)                                     (:710-711)
#   4. one sample(1:length(z_idx1), nnzidx, replace=false) per search step        (:731)
#   5. one sample(...) per perturbation replicate                                 (:772)
# so the harness seeds, runs sclens() capturing its printed lines (the number of search steps and the selected sparsity
# are only printed, :753/:762), re-seeds with the same seed and replays exactly those calls with the same arguments.
using Random, SparseArrays, DataFrames, StatsBase
using Distributions: Normal

ref_dir = ARGS[1]
out_dir = ARGS[2]
csv     = length(ARGS) >= 3 ? ARGS[3] : joinpath(ref_dir, "data", "Real_Zheng_data", "z_data_785.csv.gz")
seed    = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 785
device  = length(ARGS) >= 5 ? ARGS[5] : "gpu"
n_pert  = length(ARGS) >= 6 ? parse(Int, ARGS[6]) : 20
p_step  = 0.001

include(joinpath(ref_dir, "src", "scLENS.jl"))      # the reference, unmodified
mkpath(out_dir)
manifest = IOBuffer()

function put(name::String, a::AbstractArray)
    a = collect(a)
    open(joinpath(out_dir, name * ".bin"), "w") do io
        write(io, a)
    end
    println(manifest, name, " ", eltype(a), " ", join(size(a), "x"))     # column-major
end
put(name::String, x::Number) = put(name, [x])

ndf = scLENS.read_file(csv)
pre_df = scLENS.preprocess(ndf)
X = scLENS.df2sparr(pre_df)                          # SparseMatrixCSC{Float32,UInt32}
N, M = size(X)
nm = min(N, M)
put("X_colptr", X.colptr); put("X_rowval", X.rowval); put("X_nzval", X.nzval); put("shape", Int64[N, M])

# ---- the run
Random.seed!(seed)
log_path = joinpath(out_dir, "stdout.txt")
res = open(log_path, "w") do io
    redirect_stdout(io) do
        scLENS.sclens(pre_df; device_=device, th=60, p_step=p_step, n_perturb=n_pert)
    end
end
lines = readlines(log_path)
sel = [l for l in lines if startswith(l, "Selected perturb sparisty: ")]
p_sel = parse(Float64, split(sel[1], ": ")[2])
i_a = findfirst(l -> startswith(l, "Calculating sparsity level"), lines)
i_b = findfirst(l -> startswith(l, "Selected perturb sparisty"), lines)
d2_trace = [parse(Float64, l) for l in lines[i_a+1:i_b-1] if tryparse(Float64, l) !== nothing]   # println(ppj_[end]) :753
n_search = length(d2_trace)
p_th = parse(Float64, split([l for l in lines if startswith(l, "spth_: ")][1], ": ")[2])

# ---- the replay (same seed, same calls, same order)
Random.seed!(seed)
nz_row, nz_col, nz_val = findnz(X)
r1 = rand(UInt32(1):UInt32(N), length(nz_val)); r2 = rand(UInt32(1):UInt32(M), length(nz_val))            # :669
sample_idx = [(i, j) for (i, j) in zip(r1, r2)]
z_idset = [(i, j) for (i, j) in zip(nz_row, nz_col)]
nzz_ = setdiff(sample_idx, z_idset)
z_idx1 = [s[1] for s in nzz_]; z_idx2 = [s[2] for s in nzz_]
put("z_idx1", UInt32.(z_idx1)); put("z_idx2", UInt32.(z_idx2))
null_perm = shuffle(collect(UInt32(1):UInt32(length(nz_val))))       # shuffle(nz_val) :275 consumes the RNG by length only
ldict = countmap(nz_col)                                                                                   # :244
row_i = vcat([sample(1:N, ldict[s], replace=false) for s in keys(ldict)]...)                               # :247
put("null_perm", null_perm); put("null_rows", UInt32.(row_i)); put("null_gene_order", UInt32.(collect(keys(ldict))))
model_norm = Normal(0, sqrt(1 / nm))
p_th_replay = sum(maximum(abs.(rand(model_norm, nm))) for _ in 1:5000) / 5000                              # :710-712
put("p_th", Float64[p_th_replay, p_th])             # replayed value and the printed one: equal when the replay is in step
p_ = 0.999
for step in 1:n_search                                                                                     # :726-760
    nnzidx = Int(round((1 - p_) * M * N))
    sple = sample(UInt32(1):UInt32(lastindex(z_idx1)), nnzidx, replace=false)
    put("search_sple_$(step)", UInt32.(sple))
    global p_ -= p_step
end
for r in 1:n_pert                                                                                          # :772
    sple = sample(UInt32(1):UInt32(lastindex(z_idx1)), Int(round((1 - p_sel) * M * N)), replace=false)
    put("perturb_sple_$(r)", UInt32.(sple))
end

# ---- the outputs
put("p_sel", p_sel); put("n_search", Int64(n_search)); put("d2_trace", d2_trace); put("seed", Int64(seed))
put("L", Float64.(res[:L])); put("L_mp", Float64.(res[:L_mp])); put("lambda_c", Float64(res[:λ]))
if haskey(res, :signal_ev)
    put("signal_ev", Float64.(res[:signal_ev])); put("signal_evec", Float64.(res[:signal_evec]))
    put("sig_id", Int64.(res[:sig_id])); put("pass", Int64(res[:pass]))
    put("m_scores", Float64.(res[:robustness_scores][:m_scores])); put("sd_scores", Float64.(res[:robustness_scores][:sd_scores]))
    put("b_", Float64.(res[:robustness_scores][:b_]))
    put("pca", Float64.(Matrix(res[:pca][!, 2:end]))); put("pca_n1", Float64.(Matrix(res[:pca_n1][!, 2:end])))
    put("gene_basis", Float64.(res[:gene_basis]))
    for k in ("TGC", "mat2_mean", "mat2_std", "norm_tgc", "cent_")
        put("rec_" * k, Float64.(vec(res[:rec_vals][k])))
    end
end
open(joinpath(out_dir, "manifest.txt"), "w") do io
    write(io, String(take!(manifest)))
end
println("fixture written to ", out_dir, ": N=", N, " M=", M, " n_signal=", haskey(res, :signal_ev) ? length(res[:signal_ev]) : 0,
        " n_search=", n_search, " p_sel=", p_sel, " (device ", device, ", seed ", seed, ")")
