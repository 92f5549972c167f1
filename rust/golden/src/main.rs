//! Dumps golden vectors from the REFERENCE stack (protocol/vdf + pasta_curves 0.4 + pasta-msm 0.1 [+ nova-snark 0.8])
//! to a JSON file consumed by tests/test_ref_vectors.py.  NOT COMPILED in the build environment of this repository.
//!
//! Every byte string is the in-memory image of the Rust value (pasta_curves `repr-c`): field elements are 32 bytes
//! (four LE u64 Montgomery limbs), affine points 72 bytes, states 96 bytes -- exactly what crosses the C ABI of
//! libvdfgpu.  Inputs come from XorShiftRng::from_seed([42; 16]), the reference's TEST_SEED (src/lib.rs).
use std::{env, fs, mem};

use ff::{Field, PrimeField};
use group::{Curve, Group};
use pasta_curves::{pallas, vesta};
use rand::SeedableRng;
use rand_xorshift::XorShiftRng;
use serde_json::{json, Value};
use vdf::minroot::{MinRootVDF, PallasVDF, State, VestaVDF};
use vdf::TEST_SEED;

fn raw<T: Copy>(v: &T) -> String {
    let p = v as *const T as *const u8;
    hex::encode(unsafe { std::slice::from_raw_parts(p, mem::size_of::<T>()) })
}
fn raws<T: Copy>(v: &[T]) -> String {
    v.iter().map(raw).collect::<Vec<_>>().join("")
}

fn field_section<F: PrimeField>(rng: &mut XorShiftRng) -> Value {
    let specials = vec![F::zero(), F::one(), -F::one(), F::from(2u64), F::from(u64::MAX)];
    let mut elems: Vec<Value> = specials
        .iter()
        .map(|f| json!({"canonical_le": hex::encode(f.to_repr().as_ref()), "mont": raw(f)}))
        .collect();
    let mut ops = vec![];
    for _ in 0..16 {
        let (a, b) = (F::random(&mut *rng), F::random(&mut *rng));
        elems.push(json!({"canonical_le": hex::encode(a.to_repr().as_ref()), "mont": raw(&a)}));
        ops.push(json!({"a": raw(&a), "b": raw(&b), "add": raw(&(a + b)), "sub": raw(&(a - b)), "mul": raw(&(a * b)),
                        "square": raw(&a.square()), "neg": raw(&-a), "invert": raw(&a.invert().unwrap())}));
    }
    json!({"elements": elems, "ops": ops})
}

macro_rules! curve_section {
    ($point:ty, $affine:ty, $scalar:ty, $msm:path, $rng:expr) => {{
        let g = <$point>::generator();
        let mut pts = vec![];
        for k in [0u64, 1, 2, 3, 0xdeadbeef] {
            let p = g * <$scalar>::from(k);
            let a: $affine = p.to_affine();
            pts.push(json!({"k": k, "affine72": raw(&a), "jacobian96_of_to_curve": raw(&<$point>::from(a))}));
        }
        // MSM of 1000 random points and scalars through pasta-msm (the reference's backend for commit())
        let n = 1000;
        let points: Vec<$affine> = (0..n).map(|_| <$point>::random(&mut *$rng).to_affine()).collect();
        let mut scalars: Vec<$scalar> = (0..n).map(|_| <$scalar>::random(&mut *$rng)).collect();
        scalars[0] = <$scalar>::zero();
        scalars[1] = <$scalar>::one();
        scalars[2] = -<$scalar>::one();
        let res: $point = $msm(&points, &scalars);
        let naive = points.iter().zip(&scalars).fold(<$point>::identity(), |acc, (p, s)| acc + *p * *s);
        assert_eq!(res, naive);
        json!({"generator_affine72": raw(&g.to_affine()), "multiples": pts,
               "msm": {"n": n, "points_affine72": raws(&points), "scalars_mont": raws(&scalars),
                       "result_affine72": raw(&res.to_affine())}})
    }};
}

// MinRootVDF is generic over nova::traits::Group (reference src/minroot.rs:10, :287-290); nova must therefore be a
// direct dependency of this program as well (it is: the reference re-exports nothing).
fn minroot_section<V: MinRootVDF<G>, G: nova::traits::Group>(rng: &mut XorShiftRng) -> Value
where
    G::Scalar: PrimeField,
{
    // test_eval's inputs (src/minroot.rs:497-516): t = 10, ten random (x, y, 0) states
    let mut vdf = V::new();
    let mut chains = vec![];
    for _ in 0..10 {
        let x = State { x: G::Scalar::random(&mut *rng), y: G::Scalar::random(&mut *rng), i: G::Scalar::zero() };
        let result = vdf.eval(x, 10);
        assert!(V::check(result, 10, x));
        // State<T> has no repr(C) in the reference: dump field by field in x, y, i order
        chains.push(json!({"t": 10, "original": [raw(&x.x), raw(&x.y), raw(&x.i)].join(""),
                           "result": [raw(&result.x), raw(&result.y), raw(&result.i)].join("")}));
    }
    json!({"chains": chains})
}

fn main() {
    let out = env::args().nth(1).expect("usage: vdf-golden <output.json>");
    let mut rng = XorShiftRng::from_seed(TEST_SEED);
    let mut doc = json!({
        "generator": "rust/golden (reference stack)",
        "seed": "XorShiftRng::from_seed([42; 16])",
        "fields": {"fp": field_section::<pallas::Base>(&mut rng), "fq": field_section::<pallas::Scalar>(&mut rng)},
        "curves": {
            "pallas": curve_section!(pallas::Point, pallas::Affine, pallas::Scalar, pasta_msm::pallas, &mut rng),
            "vesta": curve_section!(vesta::Point, vesta::Affine, vesta::Scalar, pasta_msm::vesta, &mut rng),
        },
        "minroot": {"pallas": minroot_section::<PallasVDF, pallas::Point>(&mut rng),
                    "vesta": minroot_section::<VestaVDF, vesta::Point>(&mut rng)},
    });
    #[cfg(feature = "r1cs")]
    {
        doc["r1cs"] = r1cs::section(&mut rng);
    }
    fs::write(&out, serde_json::to_string_pretty(&doc).unwrap()).unwrap();
    eprintln!("wrote {out}");
}

#[cfg(feature = "r1cs")]
mod r1cs {
    //! [R] nova-snark 0.8: R1CSShape / R1CSGens / R1CSWitness / RelaxedR1CS{Witness,Instance}, ShapeCS / SatisfyingAssignment.
    use super::*;
    use bellperson::{gadgets::num::AllocatedNum, ConstraintSystem};
    use nova::{
        bellperson::{r1cs::{NovaShape, NovaWitness}, shape_cs::ShapeCS, solver::SatisfyingAssignment},
        r1cs::{R1CSInstance, R1CSWitness, RelaxedR1CSInstance, RelaxedR1CSWitness},
        traits::circuit::StepCircuit,
    };
    use vdf::nova::proof::InverseMinRootCircuit;

    type G1 = pallas::Point;

    pub fn section(rng: &mut XorShiftRng) -> Value {
        // the step circuit alone (not the augmented circuit): shape from ShapeCS, two satisfying witnesses
        let t = 5u64; // test_nova_proof's num_iters_per_step (src/nova/proof.rs:405)
        let x0 = State { x: pallas::Scalar::random(&mut *rng), y: pallas::Scalar::zero(), i: pallas::Scalar::zero() };
        let (_z0, circuits) = InverseMinRootCircuit::<G1>::eval_and_make_circuits(PallasVDF::new(), t, 2, x0);
        let mut cs: ShapeCS<G1> = ShapeCS::new();
        let z: Vec<_> = (0..3).map(|k| AllocatedNum::alloc(cs.namespace(|| format!("z{k}")), || Ok(pallas::Scalar::zero())).unwrap()).collect();
        circuits[0].synthesize(&mut cs, &z).unwrap();
        let (shape, gens) = cs.r1cs_shape();
        let mut wit = vec![];
        for c in &circuits {
            let mut cs = SatisfyingAssignment::<G1>::new();
            let r = c.result.unwrap();
            let z = vec![AllocatedNum::alloc(cs.namespace(|| "x"), || Ok(r.x)).unwrap(),
                         AllocatedNum::alloc(cs.namespace(|| "y"), || Ok(r.y)).unwrap(),
                         AllocatedNum::alloc(cs.namespace(|| "i"), || Ok(r.i)).unwrap()];
            c.synthesize(&mut cs, &z).unwrap();
            wit.push(cs.r1cs_instance_and_witness(&shape, &gens).unwrap());
        }
        let (u1, w1) = (&wit[0].0, &wit[0].1);
        let (u2, w2) = (&wit[1].0, &wit[1].1);
        let ru1 = RelaxedR1CSInstance::from_r1cs_instance(&gens, &shape, u1);
        let rw1 = RelaxedR1CSWitness::from_r1cs_witness(&shape, w1);
        let z1 = [w1.W.clone(), vec![pallas::Scalar::one()], u1.X.clone()].concat();
        let (az, bz, cz) = shape.multiply_vec(&z1).unwrap();
        let (t_vec, comm_t) = shape.commit_T(&gens, &ru1, &rw1, u2, w2).unwrap();
        let r = pallas::Scalar::from(0x1234_5678_9abc_def0u64);
        let folded = rw1.fold(w2, &t_vec, &r).unwrap();
        let coo = |m: &Vec<(usize, usize, pallas::Scalar)>| m.iter().map(|(r, c, v)| json!([r, c, raw(v)])).collect::<Vec<_>>();
        json!({"t": t, "num_cons": shape.num_cons, "num_vars": shape.num_vars, "num_io": shape.num_io,
               "A": coo(&shape.A), "B": coo(&shape.B), "C": coo(&shape.C),
               "gens_affine72": raws(&gens.gens.gens), "W1": raws(&w1.W), "X1": raws(&u1.X), "W2": raws(&w2.W), "X2": raws(&u2.X),
               "Az1": raws(&az), "Bz1": raws(&bz), "Cz1": raws(&cz), "T": raws(&t_vec),
               "comm_W1_affine72": raw(&u1.comm_W.comm.to_affine()), "comm_T_affine72": raw(&comm_t.comm.to_affine()),
               "r": raw(&r), "W_folded": raws(&folded.W), "E_folded": raws(&folded.E)})
    }
}
