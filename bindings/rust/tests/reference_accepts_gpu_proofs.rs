//! Pins the parity of the GPU engine to the REFERENCE: the documents under tests/data/ were produced by the engine
//! (tools/emit_wire_fixtures.py on a B200) in the reference's serde-JSON format; here the reference deserialises them with its
//! own `Deserialize` impls and judges them with its own verifiers.  One green run pins every encoding the Python/C oracles
//! could only recall ([R] in oracle/gg20_oracle.py): `BigInt::to_bytes`, `chain_point`, `result_scalar`, the response
//! conventions of the sigma proofs, the Paillier CRT, the serde leaf formats.  Until then parity stays "unpinned".
//!
//! Run (needs cargo + network for the crates; no GPU needed — the documents are committed):
//!     cd bindings/rust && cargo test --test reference_accepts_gpu_proofs
use curv::arithmetic::traits::*;
use curv::cryptographic_primitives::proofs::sigma_dlog::DLogProof;
use curv::elliptic::curves::{Point, Scalar, Secp256k1};
use curv::BigInt;
use multi_party_ecdsa::protocols::multi_party_ecdsa::gg_2020::party_i::{verify, KeyGenBroadcastMessage1, KeyGenDecommitMessage1, SignatureRecid};
use multi_party_ecdsa::protocols::multi_party_ecdsa::gg_2020::state_machine::keygen::LocalKey;
use multi_party_ecdsa::utilities::mta::range_proofs::AliceProof;
use multi_party_ecdsa::utilities::mta::{MessageA, MessageB};
use multi_party_ecdsa::utilities::zk_pdl_with_slack::{PDLwSlackProof, PDLwSlackStatement};
use paillier::{DecryptionKey, EncryptionKey};
use serde::Deserialize;
use sha2::Sha256;
use zk_paillier::zkproofs::DLogStatement;

fn load<T: for<'de> Deserialize<'de>>(name: &str) -> T {
    let path = format!("{}/tests/data/{}.json", env!("CARGO_MANIFEST_DIR"), name);
    serde_json::from_str(&std::fs::read_to_string(&path).unwrap_or_else(|e| panic!("{}: {}", path, e))).unwrap_or_else(|e| panic!("{}: {}", path, e))
}

#[derive(Deserialize)]
struct AliceDoc { cipher: BigInt, ek: EncryptionKey, dlog_statement: DLogStatement, proof: AliceProof }
#[test]
fn alice_proof_from_the_gpu_verifies() {
    let d: AliceDoc = load("alice_proof");
    assert!(d.proof.verify(&d.cipher, &d.ek, &d.dlog_statement)); // utilities/mta/range_proofs.rs:105
}

#[derive(Deserialize)]
struct MessageADoc { message: MessageA, ek: EncryptionKey, dlog_statements: Vec<DLogStatement> }
#[test]
fn message_a_range_proofs_verify() {
    let d: MessageADoc = load("message_a");
    assert_eq!(d.message.range_proofs.len(), d.dlog_statements.len());
    for (pf, st) in d.message.range_proofs.iter().zip(&d.dlog_statements) {
        assert!(pf.verify(&d.message.c, &d.ek, st));
    }
}

#[derive(Deserialize)]
struct MessageBDoc { message: MessageB, dk: DecryptionKey, a: Scalar<Secp256k1>, beta: Scalar<Secp256k1>, expected_alpha: Scalar<Secp256k1>,
                     expected_alpha_plus_beta: Scalar<Secp256k1> }
#[test]
fn message_b_decrypts_to_the_engines_alpha() {
    let d: MessageBDoc = load("message_b");
    let (alpha, _) = d.message.verify_proofs_get_alpha(&d.dk, &d.a).expect("MessageB rejected"); // utilities/mta/mod.rs:160
    assert_eq!(alpha, d.expected_alpha);
    assert_eq!(&alpha + &d.beta, d.expected_alpha_plus_beta); // = a * b, utilities/mta/test.rs:12-18
}

#[derive(Deserialize)]
struct DLogDoc { proof: DLogProof<Secp256k1, Sha256> }
#[test]
fn dlog_proof_verifies() {
    let d: DLogDoc = load("dlog_proof");
    DLogProof::verify(&d.proof).expect("DLogProof rejected");
}

#[derive(Deserialize)]
struct PdlDoc { statement: PDLwSlackStatement, proof: PDLwSlackProof }
#[test]
fn pdl_with_slack_proof_verifies() {
    let d: PdlDoc = load("pdl");
    d.proof.verify(&d.statement).expect("PDLwSlackProof rejected"); // utilities/zk_pdl_with_slack/mod.rs:127
}

#[derive(Deserialize)]
struct Bc1Doc { message: KeyGenBroadcastMessage1, decommit: KeyGenDecommitMessage1 }
#[test]
fn keygen_broadcast_verifies() {
    use curv::cryptographic_primitives::commitments::{hash_commitment::HashCommitment, traits::Commitment};
    let d: Bc1Doc = load("keygen_broadcast1");
    d.message.correct_key_proof.verify(&d.message.e, zk_paillier::zkproofs::SALT_STRING).expect("NiCorrectKeyProof rejected"); // gg_2020/party_i.rs:288-291
    d.message.composite_dlog_proof_base_h1.verify(&d.message.dlog_statement).expect("CompositeDLogProof(h1) rejected");
    let st2 = DLogStatement { N: d.message.dlog_statement.N.clone(), g: d.message.dlog_statement.ni.clone(), ni: d.message.dlog_statement.g.clone() };
    d.message.composite_dlog_proof_base_h2.verify(&st2).expect("CompositeDLogProof(h2) rejected");
    let com = HashCommitment::<Sha256>::create_commitment_with_user_defined_randomness(&BigInt::from_bytes(&d.decommit.y_i.to_bytes(true)), &d.decommit.blind_factor);
    assert_eq!(com, d.message.com); // gg_2020/party_i.rs:279-282
}

#[test]
fn local_key_deserialises() {
    let k: LocalKey<Secp256k1> = load("local_key"); // keygen/rounds.rs:309-322
    assert_eq!((k.i, k.t, k.n), (1, 1, 3));
    assert_eq!(k.pk_vec.len(), 3);
    assert_eq!(Point::generator() * &k.keys_linear.x_i, k.pk_vec[0]);
    assert_eq!(&k.paillier_dk.p * &k.paillier_dk.q, k.paillier_key_vec[0].n);
    assert_eq!(k.vss_scheme.commitments[0], k.y_sum_s);
}

#[derive(Deserialize)]
struct SigDoc { sig: SignatureRecid, y: Point<Secp256k1>, message: BigInt }
#[test]
fn signature_of_a_gpu_offline_stage_verifies() {
    let d: SigDoc = load("signature");
    verify(&d.sig, &d.y, &d.message).expect("ECDSA signature rejected by the reference's verify"); // gg_2020/party_i.rs:913-936
}

// ---- Lindell-2017: the engine's ECDDHProof and a whole two-party signature -----------------------
use multi_party_ecdsa::protocols::two_party_ecdsa::lindell_2017::{party_one, party_two};

#[derive(Deserialize)]
struct L17EphDoc { message: party_one::EphKeyGenFirstMsg }
#[test]
fn lindell17_ephemeral_message_from_the_gpu_verifies() {
    let d: L17EphDoc = load("lindell17_eph_first_message");
    // party_two::EphKeyGenSecondMsg::verify_and_decommit (party_two.rs:374-387) needs party two's own witness only to hand it back:
    let (_first, witness, _pair) = party_two::EphKeyGenFirstMsg::create_commitments();
    party_two::EphKeyGenSecondMsg::verify_and_decommit(witness, &d.message).expect("ECDDHProof of the engine rejected");
}

#[derive(Deserialize)]
struct L17SigDoc { party_one_ec_key: party_one::EcKeyPair, party_one_paillier: party_one::PaillierKeyPair, party_one_eph: party_one::EphEcKeyPair,
                   party_two_eph_public: Point<Secp256k1>, c3: BigInt, pubkey: Point<Secp256k1>, message: BigInt, signature: party_one::Signature, recid: u8 }
#[test]
fn lindell17_signature_is_what_the_reference_computes_from_the_gpu_partial_signature() {
    let d: L17SigDoc = load("lindell17_signature");
    let private = party_one::Party1Private::set_private_key(&d.party_one_ec_key, &d.party_one_paillier);
    // party one's side on the CPU with the reference's own Paillier decrypt (party_one.rs:486-517) over party two's GPU-made c3 ...
    let sig = party_one::Signature::compute(&private, &d.c3, &d.party_one_eph, &d.party_two_eph_public);
    assert_eq!((&sig.r, &sig.s), (&d.signature.r, &d.signature.s)); // ... equals what tecdsa_l17_sign_batch produced on the device
    let with_recid = party_one::Signature::compute_with_recid(&private, &d.c3, &d.party_one_eph, &d.party_two_eph_public);
    assert_eq!(with_recid.recid, d.recid);
    party_one::verify(&d.signature, &d.pubkey, &d.message).expect("signature rejected by party_one::verify"); // party_one.rs:567-592
}

// ---- the round messages of an offline stage produced by the engine (f2: state_machine/sign.rs:478-490) ---------------------------
use curv::cryptographic_primitives::proofs::sigma_correct_homomorphic_elgamal_enc::{HomoELGamalProof, HomoElGamalStatement};
use curv::cryptographic_primitives::proofs::sigma_valid_pedersen::PedersenProof;
use multi_party_ecdsa::protocols::multi_party_ecdsa::gg_2020::state_machine::sign::OfflineProtocolMessage;
use round_based::Msg;

#[derive(Deserialize)]
struct OfflineMsgDoc { messages: serde_json::Value, #[serde(rename = "R")] r: Point<Secp256k1>, #[serde(rename = "T_i")] t_i: Point<Secp256k1>,
                       pdl_statement: PDLwSlackStatement }
#[test]
fn offline_round_messages_of_the_engine_are_reference_messages() {
    let d: OfflineMsgDoc = load("offline_messages");
    // shape: the whole array is what a `gg20_sm_manager` room would carry
    let typed: Vec<Msg<OfflineProtocolMessage>> = serde_json::from_value(d.messages.clone()).expect("not Vec<Msg<OfflineProtocolMessage>>");
    assert_eq!(serde_json::to_value(&typed).unwrap(), d.messages, "re-serialisation differs: a leaf encoding is not the reference's");
    let body = |kind: &str| d.messages.as_array().unwrap().iter().find_map(|m| m["body"].get(kind).cloned()).unwrap_or_else(|| panic!("no {}", kind));
    // M3 = (DeltaI, TI, TIProof): the Pedersen proof of T_i (sign/rounds.rs:371-378)
    let m3 = body("M3");
    let ped: PedersenProof<Secp256k1, Sha256> = serde_json::from_value(m3[2].clone()).unwrap();
    assert_eq!(ped.com, d.t_i);
    PedersenProof::verify(&ped).expect("PedersenProof of the engine rejected");
    // M5 = (RDash, Vec<PDLwSlackProof>) (party_i.rs:719-766)
    let m5 = body("M5");
    let pdl: Vec<PDLwSlackProof> = serde_json::from_value(m5[1].clone()).unwrap();
    pdl[0].verify(&d.pdl_statement).expect("PDLwSlackProof of the engine rejected");
    // M6 = (SI, HEGProof) (party_i.rs:801-833)
    let m6 = body("M6");
    let s_i: Point<Secp256k1> = serde_json::from_value(m6[0].clone()).unwrap();
    let heg: HomoELGamalProof<Secp256k1, Sha256> = serde_json::from_value(m6[1].clone()).unwrap();
    let st = HomoElGamalStatement { G: d.r.clone(), H: Point::<Secp256k1>::base_point2().clone(), Y: Point::generator().to_point(), D: d.t_i.clone(), E: s_i };
    heg.verify(&st).expect("HomoELGamalProof of the engine rejected");
}
