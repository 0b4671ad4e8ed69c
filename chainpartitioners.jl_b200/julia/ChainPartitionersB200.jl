# ChainPartitionersB200.jl -- the Julia-side binding a ChainPartitioners.jl maintainer would add
# (one more glue file next to src/glue_*.jl, or a same-named package): it keeps the reference's API
#   partition_stripe(A, K, method, args...; kwargs...)      (src/DynamicSplitter.jl:15,
#                                                            src/BisectCostBottleneckSplitter.jl:6,
#                                                            src/LazyBisectCostBottleneckSplitter.jl:8,140,260)
#   pack_stripe(A, method, args...; kwargs...)              (src/DynamicChunker.jl:20, src/ConvexTotalChunker.jl:141,
#                                                            src/OverlapChunker.jl:6, src/StrictChunker.jl:5)
#   oracle_stripe / bound_stripe                            (src/Costs.jl:3-19)
# and forwards to libchainb200.so (include/chainb200.h) with `ccall`.  Julia is not installed in the
# build image, so this file is NOT executed by the test-suite; it is deliberately mechanical: marshal
# colptr/rowval/coefficients, wrap the returned split vector in SplitPartition.  The Python mirror in
# chainpartitioners.jl_b200/api.py exercises exactly the same C entry points and is what the tests run.
module ChainPartitionersB200

using SparseArrays
using ChainPartitioners
import ChainPartitioners: partition_stripe, pack_stripe, oracle_stripe, bound_stripe, dominancecount, dominancesum, rookcount!, rooksum!,
    AffineWorkModel, AffineConnectivityModel, AffineMonotonizedSymmetricConnectivityModel,
    AffineSymmetricConnectivityModel, AffineHyperedgeCutModel, AffineSymmetricEdgeCutModel, AffineEnvelopeModel,
    AffinePrimaryConnectivityModel, AffineSecondaryConnectivityModel, AffinePrimaryEdgeCutModel, AffineSecondaryEdgeCutModel,
    ColumnBlockComponentCostModel, BlockComponentCostModel, block_component,
    ConstrainedCost, VertexCount, FeasibleCost, SplitPartition,
    DynamicBottleneckSplitter, DynamicTotalSplitter, BisectCostBottleneckSplitter, BisectIndexBottleneckSplitter, LazyBisectCostBottleneckSplitter,
    FlipBisectCostBottleneckSplitter, LazyFlipBisectCostBottleneckSplitter, FlipBisectIndexBottleneckSplitter,
    DynamicBottleneckChunker, DynamicTotalChunker, ConvexTotalChunker, ConvexTotalSplitter, OverlapChunker, StrictChunker, EquiChunker, EquiSplitter

const lib = get(ENV, "CHAINB200_LIB", "libchainb200.so")

# ---- C structs (binary layout of cpb_model / cpb_constraint) ---------------------------------
struct CModel
    kind::Int32
    is_float::Int32
    coef::NTuple{8, Float64}
    R::Int32
    w_tab::Int32
    u_tab::Int32
    _pad::Int32
    alpha_col::Ptr{Float64}
    beta_col::Ptr{Float64}
    beta_row::Ptr{Float64}
end

struct CConstraint
    enabled::Int32
    _pad::Int32
    w_coef::NTuple{3, Int64}
    w_max::Int64
end

struct CpbError <: Exception
    code::Int
    msg::String
end

check(rc) = rc == 0 ? nothing : throw(CpbError(rc, unsafe_string(ccall((:cpb_last_error, lib), Cstring, ()))))

pad8(xs...) = ntuple(i -> i <= length(xs) ? Float64(xs[i]) : 0.0, 8)
isfloat(::Type{<:AbstractFloat}) = Int32(1)
isfloat(::Type) = Int32(0)
affine(kind, ::Type{Tv}, xs...) where {Tv} = (CModel(kind, isfloat(Tv), pad8(xs...), 0, 0, 0, 0, C_NULL, C_NULL, C_NULL), Any[])

cmodel(m::AffineWorkModel{Tv}, args...) where {Tv} = affine(0, Tv, m.α, m.β_vertex, m.β_pin)
cmodel(m::AffineConnectivityModel{Tv}, args...) where {Tv} = affine(1, Tv, m.α, m.β_vertex, m.β_pin, m.β_net)
cmodel(m::AffineMonotonizedSymmetricConnectivityModel{Tv}, args...) where {Tv} = affine(2, Tv, m.α, m.β_vertex, m.β_over_pin, m.β_dia_net, m.Δ_pins)
cmodel(m::AffineSymmetricConnectivityModel{Tv}, args...) where {Tv} = affine(3, Tv, m.α, m.β_vertex, m.β_pin, m.β_local_net, m.β_remote_net)
cmodel(m::AffineHyperedgeCutModel{Tv}, args...) where {Tv} = affine(4, Tv, m.α, m.β_vertex, m.β_pin, m.β_self_net, m.β_cut_net)
cmodel(m::AffineSymmetricEdgeCutModel{Tv}, args...) where {Tv} = affine(5, Tv, m.α, m.β_vertex, m.β_self_pin, m.β_cut_pin)
cmodel(m::AffineEnvelopeModel{Tv}, args...) where {Tv} = affine(6, Tv, m.α, m.β_vertex, m.β_pin, m.β_net)
cmodel(m::AffinePrimaryConnectivityModel{Tv}, args...) where {Tv} = affine(9, Tv, m.α, m.β_vertex, m.β_pin, m.β_local_net, m.β_remote_net)  # needs Π
cmodel(m::AffineSecondaryConnectivityModel{Tv}, args...) where {Tv} = affine(10, Tv, m.α, m.β_vertex, m.β_pin, m.β_local_net, m.β_remote_net)  # needs Π
cmodel(m::AffinePrimaryEdgeCutModel{Tv}, args...) where {Tv} = affine(11, Tv, m.α, m.β_vertex, m.β_self_pin, m.β_cut_pin)  # needs Π
cmodel(m::AffineSecondaryEdgeCutModel{Tv}, args...) where {Tv} = affine(12, Tv, m.α, m.β_vertex, m.β_self_pin, m.β_cut_pin)  # needs Π

# Functors cannot cross the ABI: tabulate block_component(f, w) (src/BlockCosts.jl:41-44) for w = 0..w_tab
tab(f, hi) = Float64[w == 0 && !(f isa Function || f isa Number) ? 0.0 : block_component(f, w) for w in 0:hi]
function cmodel(m::ColumnBlockComponentCostModel{Tv}, w_tab, u_tab) where {Tv}
    a, b = tab(m.α_col, w_tab), tab(m.β_col, w_tab)
    (CModel(7, isfloat(Tv), pad8(), 1, w_tab, 0, 0, pointer(a), pointer(b), C_NULL), Any[a, b])
end
function cmodel(m::BlockComponentCostModel{Tv, R}, w_tab, u_tab) where {Tv, R}
    a = tab(m.α_col, w_tab)
    bc = reduce(vcat, (tab(f, w_tab) for f in m.β_col); init = Float64[])
    br = reduce(vcat, (tab(f, u_tab) for f in m.β_row); init = Float64[])
    (CModel(8, isfloat(Tv), pad8(), R, w_tab, u_tab, 0, pointer(a), pointer(bc), pointer(br)), Any[a, bc, br])
end

constraint(::Any) = CConstraint(0, 0, (0, 0, 0), 0)
constraint(c::ConstrainedCost{<:Any, VertexCount}) = CConstraint(1, 0, (0, 1, 0), Int64(c.w_max))
constraint(c::ConstrainedCost{<:Any, <:AffineWorkModel{<:Integer}}) = CConstraint(1, 0, (Int64(c.w.α), Int64(c.w.β_vertex), Int64(c.w.β_pin)), Int64(c.w_max))
constraint(c::ConstrainedCost{<:Any, FeasibleCost}) = CConstraint(0, 0, (0, 0, 0), 0)
basecost(f) = f
basecost(c::ConstrainedCost) = c.f

# ---- handles ---------------------------------------------------------------------------------
mutable struct DeviceMatrix
    h::Ptr{Cvoid}
    m::Int
    n::Int
    function DeviceMatrix(A::SparseMatrixCSC{Tv, Int64}) where {Tv}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        (m, n) = size(A)
        GC.@preserve A check(ccall((:cpb_matrix_create, lib), Cint,
            (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}), m, n, nnz(A), A.colptr, A.rowval, h))
        finalizer(x -> ccall((:cpb_matrix_destroy, lib), Cvoid, (Ptr{Cvoid},), x.h), new(h[], m, n))
    end
    function DeviceMatrix(A::SparseMatrixCSC{Tv, Int32}) where {Tv}  # half the upload
        h = Ref{Ptr{Cvoid}}(C_NULL)
        (m, n) = size(A)
        GC.@preserve A check(ccall((:cpb_matrix_create_i32, lib), Cint,
            (Int64, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}), m, n, nnz(A), A.colptr, A.rowval, h))
        finalizer(x -> ccall((:cpb_matrix_destroy, lib), Cvoid, (Ptr{Cvoid},), x.h), new(h[], m, n))
    end
end

mutable struct DeviceOracle
    h::Ptr{Cvoid}
    A::DeviceMatrix
    function DeviceOracle(A::DeviceMatrix, f, Π = nothing; w_tab = A.n)
        u_tab = Π === nothing ? 0 : maximum(diff(Π.spl))
        (cm, keep) = cmodel(basecost(f), w_tab, u_tab)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        spl = Π === nothing ? C_NULL : pointer(Π.spl)
        K = Π === nothing ? 0 : Π.K
        GC.@preserve keep Π check(ccall((:cpb_oracle_create, lib), Cint,
            (Ptr{Cvoid}, Ref{CModel}, Ptr{Int64}, Int64, Ref{Ptr{Cvoid}}), A.h, cm, spl, K, h))
        finalizer(x -> ccall((:cpb_oracle_destroy, lib), Cvoid, (Ptr{Cvoid},), x.h), new(h[], A))
    end
end

# ocl(j, j′[, k]) -- src/ConnectivityCosts.jl:58-64 and friends
function (ocl::DeviceOracle)(j::Integer, j′::Integer, k...)
    out = Ref{Float64}(0.0)
    check(ccall((:cpb_oracle_query, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ptr{Int64}, Ref{Float64}),
                ocl.h, 1, Int64(j), Int64(j′), C_NULL, out))
    return out[]
end
function query(ocl::DeviceOracle, j::Vector{Int64}, j′::Vector{Int64})
    out = Vector{Float64}(undef, length(j))
    check(ccall((:cpb_oracle_query, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                ocl.h, length(j), j, j′, C_NULL, out))
    return out
end

function bound_stripe(ocl::DeviceOracle, K)
    out = Vector{Float64}(undef, 2)
    check(ccall((:cpb_bound_stripe, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), ocl.h, K, out))
    return (out[1], out[2])
end

# ---- partition_stripe / pack_stripe -----------------------------------------------------------
struct OnB200{Mtd}   # partition_stripe(A, K, OnB200(DynamicBottleneckSplitter(f))) etc.
    mtd::Mtd
end

split_code(::DynamicBottleneckSplitter) = (0, 0.0)
split_code(::DynamicTotalSplitter) = (1, 0.0)
split_code(m::BisectCostBottleneckSplitter) = (2, Float64(m.ϵ))
split_code(m::LazyBisectCostBottleneckSplitter) = (3, Float64(m.ϵ))
split_code(::ConvexTotalSplitter) = (8, 0.0)
split_code(::BisectIndexBottleneckSplitter) = (12, 0.0)
split_code(m::FlipBisectCostBottleneckSplitter) = (6, Float64(m.ϵ))          # Flip family: the secondary (decreasing) models
split_code(m::LazyFlipBisectCostBottleneckSplitter) = (7, Float64(m.ϵ))
split_code(::FlipBisectIndexBottleneckSplitter) = (13, 0.0)
split_code(::DynamicBottleneckChunker) = (10, 0.0)   # partition_stripe(A, K, ::AbstractDynamicChunker), DynamicSplitter.jl:52-87
split_code(::DynamicTotalChunker) = (11, 0.0)

function partition_stripe(A::SparseMatrixCSC{Tv, Int64}, K, method::OnB200, args...; kwargs...) where {Tv}
    dA = DeviceMatrix(A)
    (code, ϵ) = split_code(method.mtd)
    ocl = DeviceOracle(dA, method.mtd.f, args...)
    spl = Vector{Int64}(undef, K + 1)
    check(ccall((:cpb_partition_stripe, lib), Cint, (Ptr{Cvoid}, Cint, Ref{CConstraint}, Float64, Int64, Ptr{Int64}),
                ocl.h, code, constraint(method.mtd.f), ϵ, K, spl))
    return SplitPartition{Int64}(K, spl)
end

pack_code(m::DynamicTotalChunker) = (0, m.f, 0.0, 0)
pack_code(m::ConvexTotalChunker) = (1, m.f, 0.0, 0)
pack_code(m::OverlapChunker) = (3, nothing, m.ρ, m.w_max)
pack_code(m::StrictChunker) = (4, nothing, 0.0, m.w_max)
pack_code(m::EquiChunker) = (5, nothing, 0.0, m.w)

function pack_stripe(A::SparseMatrixCSC{Tv, Int64}, method::OnB200, args...; n_nets = nothing, kwargs...) where {Tv}
    (m, n) = size(A)
    dA = DeviceMatrix(A)
    (code, f, ρ, w_max) = pack_code(method.mtd)
    con = constraint(f)
    ocl = f === nothing ? nothing : DeviceOracle(dA, f, args...; w_tab = con.enabled == 1 ? min(n, con.w_max ÷ max(con.w_coef[2], 1) + 1) : n)
    spl = Vector{Int64}(undef, n + 1)
    nn = Vector{Int64}(undef, max(n, 1))
    K = Ref{Int64}(0)
    check(ccall((:cpb_pack_stripe, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{CConstraint}, Float64, Int64, Ptr{Int64}, Ref{Int64}, Ptr{Int64}),
                dA.h, ocl === nothing ? C_NULL : ocl.h, code, con, ρ, w_max, spl, K, nn))
    if n_nets isa Ref
        n_nets[] = nn[1:K[]]
    end
    return SplitPartition{Int64}(K[], resize!(spl, K[] + 1))
end

# ---- SparsePrefixMatrices.jl on the device: dominancecount / dominancesum / rookcount! / rooksum! --------------------
# (src/SparsePrefixMatrices.jl:31-58, 440-460, 840-851, 1056-1063).  One handle type; `C[i, j]` as in the reference.
# The device takes the place of the hint argument: `dominancecount(OnB200(nothing), A)`.
mutable struct DevicePrefixMatrix{Tv} <: AbstractMatrix{Tv}
    h::Ptr{Cvoid}
    m::Int
    n::Int
    summed::Bool
    function DevicePrefixMatrix{Tv}(m, n, N, pos, idx::Vector{Int64}, val) where {Tv}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve pos idx val check(ccall((:cpb_prefix_create, lib), Cint,
            (Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}), m, n, N,
            pos === nothing ? C_NULL : pointer(pos), idx, val === nothing ? C_NULL : Ptr{Int64}(pointer(val)), h))
        finalizer(x -> ccall((:cpb_prefix_destroy, lib), Cvoid, (Ptr{Cvoid},), x.h), new{Tv}(h[], m, n, val !== nothing))
    end
end
Base.size(P::DevicePrefixMatrix) = (P.m + 1, P.n + 1)
function Base.getindex(P::DevicePrefixMatrix{Tv}, i::Integer, j::Integer) where {Tv}
    out = Ref{Int64}(0)
    check(ccall((:cpb_prefix_query, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}),
                P.h, 1, Int64(i), Int64(j), P.summed ? C_NULL : out, P.summed ? out : C_NULL))
    return reinterpret(Tv, out[])
end
dominancecount(::OnB200, A::SparseMatrixCSC{Tv, Int64}; kwargs...) where {Tv} =
    DevicePrefixMatrix{Int64}(size(A)..., nnz(A), A.colptr, A.rowval, nothing)
dominancesum(::OnB200, A::SparseMatrixCSC{Tv, Int64}; kwargs...) where {Tv <: Union{Int64, UInt64}} =
    DevicePrefixMatrix{Tv}(size(A)..., nnz(A), A.colptr, A.rowval, A.nzval)
rookcount!(::OnB200, N, idx::Vector{Int64}; kwargs...) = DevicePrefixMatrix{Int64}(N, N, N, nothing, idx, nothing)
rooksum!(::OnB200, N, idx::Vector{Int64}, val::Vector{Tv}; kwargs...) where {Tv <: Union{Int64, UInt64}} =
    DevicePrefixMatrix{Tv}(N, N, N, nothing, idx, val)

# partition_plaid / pack_plaid need no glue: AlternatingPartitioner(OnB200(mtd1), OnB200(mtd2)) already
# alternates partition_stripe calls on A and adjointpattern(A) (src/AlternatingPartitioner.jl:18-32).

end # module
