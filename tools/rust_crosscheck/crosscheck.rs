//! Cross-check dumper: runs the REFERENCE's own code on inputs exported by `tools/rust_crosscheck/export_inputs.py` of
//! the B200 repository and writes its outputs for `compare.py`.  Drop into `plonky2/src/crosscheck.rs` (see README.md).
//! Never compiled in the B200 repository (no Rust toolchain there).
#![cfg(test)]
use std::fs;
use std::path::PathBuf;

use crate::field::extension::quadratic::QuadraticExtension;
use crate::field::goldilocks_field::GoldilocksField;
use crate::field::polynomial::PolynomialValues;
use crate::field::types::{Field, PrimeField64};
use crate::fri::oracle::PolynomialBatch;
use crate::fri::structure::{FriBatchInfo, FriInstanceInfo, FriOracleInfo, FriPolynomialInfo};
use crate::fri::{FriConfig, FriParams};
use crate::gates::arithmetic_extension::ArithmeticExtensionGate;
use crate::gates::exponentiation::ExponentiationGate;
use crate::gates::gate::Gate;
use crate::gates::high_degree_interpolation::HighDegreeInterpolationGate;
use crate::gates::low_degree_interpolation::LowDegreeInterpolationGate;
use crate::gates::multiplication_extension::MulExtensionGate;
use crate::gates::poseidon_mds::PoseidonMdsGate;
use crate::gates::reducing::ReducingGate;
use crate::gates::reducing_extension::ReducingExtensionGate;
use crate::hash::hash_types::HashOut;
use crate::iop::challenger::Challenger;
use crate::plonk::config::{GenericConfig, PoseidonGoldilocksConfig};
use crate::plonk::vars::{EvaluationVarsBaseBatch};
use crate::util::timing::TimingTree;

const D: usize = 2;
type C = PoseidonGoldilocksConfig;
type F = <C as GenericConfig<D>>::F;
type FE = QuadraticExtension<GoldilocksField>;

fn dir(var: &str) -> PathBuf { PathBuf::from(std::env::var(var).expect("set P2X_IN / P2X_OUT")) }
fn read_u64(name: &str) -> Vec<u64> {
    let b = fs::read(dir("P2X_IN").join(name)).unwrap_or_else(|_| panic!("missing input {}", name));
    b.chunks_exact(8).map(|c| u64::from_le_bytes(c.try_into().unwrap())).collect()
}
fn write_u64(name: &str, v: &[u64]) {
    fs::create_dir_all(dir("P2X_OUT")).unwrap();
    let mut b = Vec::with_capacity(v.len() * 8);
    for x in v { b.extend_from_slice(&x.to_le_bytes()); }
    fs::write(dir("P2X_OUT").join(name), b).unwrap();
}
fn f(x: u64) -> F { F::from_canonical_u64(x) }
fn u(x: F) -> u64 { x.to_canonical_u64() }
fn hashes(h: &[HashOut<F>]) -> Vec<u64> { h.iter().flat_map(|h| h.elements.iter().map(|&e| u(e))).collect() }

/// `commit_<k>.meta = [n_log, polys, rate_bits, cap_height]`, `commit_<k>.values = [polys][n]`
/// PolynomialBatch::from_values, plonky2/src/fri/oracle.rs:709-731
#[test]
fn crosscheck_commits() {
    for k in 0.. {
        let stem = format!("commit_{}", k);
        if !dir("P2X_IN").join(format!("{}.meta.bin", stem)).exists() { break; }
        let m = read_u64(&format!("{}.meta.bin", stem));
        let (n, polys, rate_bits, cap_height) = (1usize << m[0], m[1] as usize, m[2] as usize, m[3] as usize);
        let vals = read_u64(&format!("{}.values.bin", stem));
        let values: Vec<PolynomialValues<F>> = (0..polys)
            .map(|c| PolynomialValues::new(vals[c * n..(c + 1) * n].iter().map(|&x| f(x)).collect()))
            .collect();
        let batch = PolynomialBatch::<F, C, D>::from_values(values, rate_bits, false, cap_height, &mut TimingTree::default(), None);
        write_u64(&format!("{}.coeffs.bin", stem), &batch.polynomials.iter().flat_map(|p| p.coeffs.iter().map(|&c| u(c))).collect::<Vec<_>>());
        write_u64(&format!("{}.leaves.bin", stem), &batch.merkle_tree.leaves.iter().flat_map(|l| l.iter().map(|&c| u(c))).collect::<Vec<_>>());
        write_u64(&format!("{}.digests.bin", stem), &hashes(&batch.merkle_tree.digests));
        write_u64(&format!("{}.cap.bin", stem), &hashes(&batch.merkle_tree.cap.0));
    }
}

/// `fri_<k>.meta = [degree_bits, rate_bits, cap_height, pow_bits, queries, n_arity, arity..., n_oracles, polys...]`,
/// `fri_<k>.oracle<j>.values`, `fri_<k>.zeta = [2]`, `fri_<k>.observed = [..]` (transcript prefix)
/// PolynomialBatch::prove_openings, plonky2/src/fri/oracle.rs:1046-1110; instance shape of circuit_data.rs:351-371
#[test]
fn crosscheck_fri() {
    for k in 0.. {
        let stem = format!("fri_{}", k);
        if !dir("P2X_IN").join(format!("{}.meta.bin", stem)).exists() { break; }
        let m = read_u64(&format!("{}.meta.bin", stem));
        let (degree_bits, rate_bits, cap_height, pow_bits, queries) = (m[0] as usize, m[1] as usize, m[2] as usize, m[3] as u32, m[4] as usize);
        let n_ar = m[5] as usize;
        let arity: Vec<usize> = m[6..6 + n_ar].iter().map(|&x| x as usize).collect();
        let n_or = m[6 + n_ar] as usize;
        let polys: Vec<usize> = m[7 + n_ar..7 + n_ar + n_or].iter().map(|&x| x as usize).collect();
        let n = 1usize << degree_bits;
        let mut timing = TimingTree::default();
        let oracles: Vec<PolynomialBatch<F, C, D>> = (0..n_or).map(|j| {
            let v = read_u64(&format!("{}.oracle{}.values.bin", stem, j));
            let values = (0..polys[j]).map(|c| PolynomialValues::new(v[c * n..(c + 1) * n].iter().map(|&x| f(x)).collect())).collect();
            PolynomialBatch::from_values(values, rate_bits, false, cap_height, &mut timing, None)
        }).collect();
        let z = read_u64(&format!("{}.zeta.bin", stem));
        let zeta = FE::from_basefield_array([f(z[0]), f(z[1])]);
        let g = F::primitive_root_of_unity(degree_bits);
        let all: Vec<FriPolynomialInfo> = (0..n_or).flat_map(|o| FriPolynomialInfo::from_range(o, 0..polys[o])).collect();
        let zs = FriPolynomialInfo::from_range(2, 0..polys[2]);
        let instance = FriInstanceInfo {
            oracles: (0..n_or).map(|_| FriOracleInfo { blinding: false }).collect(),
            batches: vec![FriBatchInfo { point: zeta, polynomials: all }, FriBatchInfo { point: zeta.scalar_mul(g), polynomials: zs }],
        };
        let params = FriParams {
            config: FriConfig { rate_bits, cap_height, proof_of_work_bits: pow_bits, reduction_strategy: crate::fri::reduction_strategies::FriReductionStrategy::Fixed(arity.clone()), num_query_rounds: queries },
            hiding: false, degree_bits, reduction_arity_bits: arity,
        };
        let mut challenger = Challenger::<F, <C as GenericConfig<D>>::Hasher>::new();
        challenger.observe_elements(&read_u64(&format!("{}.observed.bin", stem)).iter().map(|&x| f(x)).collect::<Vec<_>>());
        let proof = PolynomialBatch::prove_openings(&instance, &oracles.iter().collect::<Vec<_>>(), &mut challenger, &params, &mut timing);
        for (i, cap) in proof.commit_phase_merkle_caps.iter().enumerate() { write_u64(&format!("{}.cap{}.bin", stem, i), &hashes(&cap.0)); }
        write_u64(&format!("{}.final_poly.bin", stem), &proof.final_poly.coeffs.iter().flat_map(|c| c.to_basefield_array().map(u)).collect::<Vec<_>>());
        write_u64(&format!("{}.pow_witness.bin", stem), &[u(proof.pow_witness)]);
        let mut rows = vec![];
        for q in &proof.query_round_proofs {
            for (leaf, path) in &q.initial_trees_proof.evals_proofs {
                rows.extend(leaf.iter().map(|&x| u(x)));
                rows.extend(hashes(&path.siblings));
            }
            for step in &q.steps {
                rows.extend(step.evals.iter().flat_map(|e| e.to_basefield_array().map(u)));
                rows.extend(hashes(&step.merkle_proof.siblings));
            }
        }
        write_u64(&format!("{}.queries.bin", stem), &rows);
        write_u64(&format!("{}.challenger_after.bin", stem), &challenger.compact().iter().map(|&x| u(x)).collect::<Vec<_>>());
    }
}

/// `perm_<k>.meta = [degree_bits, num_routed, num_challenges, max_degree]`, `.wires = [routed][n]`, `.sigmas = [routed][n]`
/// (values of the sigma polynomials on H), `.k_is`, `.betas`, `.gammas`.
/// Restates the body of all_wires_permutation_partial_products / wires_permutation_partial_products_and_zs
/// (plonk/prover.rs:702-786) with the public(crate) pieces it is made of, so no witness/prover-data plumbing is needed.
#[test]
fn crosscheck_partial_products() {
    use crate::util::partial_products::{partial_products_and_z_gx, quotient_chunk_products};
    for k in 0.. {
        let stem = format!("perm_{}", k);
        if !dir("P2X_IN").join(format!("{}.meta.bin", stem)).exists() { break; }
        let m = read_u64(&format!("{}.meta.bin", stem));
        let (n, routed, nch, max_degree) = (1usize << m[0], m[1] as usize, m[2] as usize, m[3] as usize);
        let wires = read_u64(&format!("{}.wires.bin", stem));
        let sigmas = read_u64(&format!("{}.sigmas.bin", stem));
        let k_is = read_u64(&format!("{}.k_is.bin", stem));
        let (betas, gammas) = (read_u64(&format!("{}.betas.bin", stem)), read_u64(&format!("{}.gammas.bin", stem)));
        let subgroup = F::two_adic_subgroup(m[0] as usize);
        let mut out = vec![];
        for c in 0..nch {
            let (beta, gamma) = (f(betas[c]), f(gammas[c]));
            let mut z_x = F::ONE;
            let mut all = vec![vec![]; n];
            for i in 0..n {
                let x = subgroup[i];
                let num: Vec<F> = (0..routed).map(|j| f(wires[j * n + i]) + beta * f(k_is[j]) * x + gamma).collect();
                let den: Vec<F> = (0..routed).map(|j| f(wires[j * n + i]) + beta * f(sigmas[j * n + i]) + gamma).collect();
                let den_inv = F::batch_multiplicative_inverse(&den);
                let q: Vec<F> = num.iter().zip(den_inv).map(|(&a, b)| a * b).collect();
                let chunks = quotient_chunk_products(&q, max_degree);
                let mut pp = partial_products_and_z_gx(z_x, &chunks);
                pp.insert(0, z_x);                      // [Z(x), partial products.., Z(gx)]
                z_x = pp.pop().unwrap();
                all[i] = pp;
            }
            for col in 0..all[0].len() { out.extend((0..n).map(|i| u(all[i][col]))); }   // polynomial-major, Z first
        }
        write_u64(&format!("{}.zs_pp.bin", stem), &out);
    }
}

fn dump_gate<G: Gate<F, D>>(name: &str, gate: G) {
    let stem = format!("gate_{}", name);
    if !dir("P2X_IN").join(format!("{}.meta.bin", stem)).exists() { return; }
    let m = read_u64(&format!("{}.meta.bin", stem));
    let (rows, nw, nc) = (m[0] as usize, m[1] as usize, m[2] as usize);
    // EvaluationVarsBaseBatch wants wire-major storage: [wire][row] (plonk/vars.rs)
    let w = read_u64(&format!("{}.wires.bin", stem));
    let k = read_u64(&format!("{}.consts.bin", stem));
    let pih = read_u64(&format!("{}.pih.bin", stem));
    let wires: Vec<F> = (0..nw).flat_map(|j| (0..rows).map(move |r| (r, j))).map(|(r, j)| f(w[r * nw + j])).collect();
    let consts: Vec<F> = (0..nc).flat_map(|j| (0..rows).map(move |r| (r, j))).map(|(r, j)| f(k[r * nc + j])).collect();
    let hash = HashOut { elements: [f(pih[0]), f(pih[1]), f(pih[2]), f(pih[3])] };
    let vars = EvaluationVarsBaseBatch::new(rows, &consts, &wires, &hash);
    let res = gate.eval_unfiltered_base_batch(vars);      // constraint-major: [constraint][row] (gates/gate.rs:73-90)
    write_u64(&format!("{}.constraints.bin", stem), &res.iter().map(|&x| u(x)).collect::<Vec<_>>());
}

/// The eight gates a recursive-verifier circuit adds, with the parameters of tests/quotient_fixtures.recursion_gate_set()
#[test]
fn crosscheck_recursion_gates() {
    let cfg = crate::plonk::circuit_data::CircuitConfig::standard_recursion_config();
    dump_gate("arithmetic_extension", ArithmeticExtensionGate::<D>::new_from_config(&cfg));
    dump_gate("mul_extension", MulExtensionGate::<D>::new_from_config(&cfg));
    dump_gate("reducing", ReducingGate::<D>::new(43));
    dump_gate("reducing_extension", ReducingExtensionGate::<D>::new(32));
    dump_gate("exponentiation", ExponentiationGate::<F, D>::new_from_config(&cfg));
    dump_gate("poseidon_mds", PoseidonMdsGate::<F, D>::new());
    dump_gate("high_degree_interpolation", HighDegreeInterpolationGate::<F, D>::new(2));
    dump_gate("low_degree_interpolation", LowDegreeInterpolationGate::<F, D>::new(4));
}
