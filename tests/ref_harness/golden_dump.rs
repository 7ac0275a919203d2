// golden_dump.rs — pins the resource-fit / quantity arithmetic of the REAL reference crate.
//
// NOT COMPILED IN THIS REPOSITORY (the build image has no rustc/cargo; DESIGN.md "Oracle and parity status").
// For a maintainer with a Rust toolchain:
//   1. cp tests/ref_harness/golden_dump.rs  <reference>/src/predicates/golden_dump.rs
//   2. append to <reference>/src/predicates.rs:      #[cfg(test)] mod golden_dump;
//   3. cd <reference> && cargo test golden_dump -- --nocapture | grep '^GOLDEN ' | sed 's/^GOLDEN //' > /tmp/ref_golden.json
//   4. python tests/ref_harness/compare.py /tmp/ref_golden.json
// Step 4 compares what the crate computed with tests/golden/gv1.json (hand-derived) and with
// tests/golden/quantity_kats.json; when it prints "PINNED" the oracle's resource_fits restatement is no longer
// "parity unpinned".
//
// It only uses what the reference itself uses on the path: PodResources::new / SubAssign (src/util.rs:17-36),
// total_pod_resources (src/util.rs:54-75), ParsedQuantity TryFrom / += / -= / <= (src/predicates.rs:28-42).
// can_pod_fit itself is async and LISTs pods through a kube Client, so its arithmetic (src/predicates.rs:27-42)
// is restated here line by line with the LIST replaced by an in-memory slice of the bound pods.
use std::collections::BTreeMap;

use k8s_openapi::api::core::v1 as corev1;
use k8s_openapi::apimachinery::pkg::api::resource::Quantity;
use kube_quantity::ParsedQuantity;

use crate::util::{
    total_pod_resources,
    PodResources,
};

fn container(req: Option<&[(&str, &str)]>) -> corev1::Container {
    let mut c = corev1::Container::default();
    c.name = "c".to_string();
    if let Some(kvs) = req {
        let mut m = BTreeMap::new();
        for (k, v) in kvs {
            m.insert(k.to_string(), Quantity(v.to_string()));
        }
        c.resources = Some(corev1::ResourceRequirements { requests: Some(m), ..Default::default() });
    }
    return c;
}

fn pod(name: &str, node_name: Option<&str>, containers: Vec<corev1::Container>) -> corev1::Pod {
    let mut p = corev1::Pod::default();
    p.metadata.namespace = Some("default".to_string());
    p.metadata.name = Some(name.to_string());
    let mut spec = corev1::PodSpec::default();
    spec.node_name = node_name.map(|s| s.to_string());
    spec.containers = containers;
    p.spec = Some(spec);
    return p;
}

fn node(name: &str, allocatable: Option<(&str, &str)>) -> corev1::Node {
    let mut n = corev1::Node::default();
    n.metadata.name = Some(name.to_string());
    if let Some((cpu, mem)) = allocatable {
        let mut m = BTreeMap::new();
        m.insert("cpu".to_string(), Quantity(cpu.to_string()));
        m.insert("memory".to_string(), Quantity(mem.to_string()));
        n.status = Some(corev1::NodeStatus { allocatable: Some(m), ..Default::default() });
    }
    return n;
}

// src/predicates.rs:27-42 with `pods_on_node` handed in instead of LISTed
fn available(node: &corev1::Node, pods_on_node: &[&corev1::Pod]) -> PodResources {
    let mut available_resources = PodResources::new();
    if let Some(corev1::NodeStatus { allocatable: Some(allocatable), .. }) = &node.status {
        available_resources.cpu = allocatable["cpu"].clone().try_into().expect("invalid node spec: allocatable cpu");
        available_resources.memory =
            allocatable["memory"].clone().try_into().expect("invalid node spec: allocatable memory");
    }
    for p in pods_on_node {
        available_resources -= total_pod_resources(p)
    }
    return available_resources;
}

fn fits(pod: &corev1::Pod, avail: &PodResources) -> bool {
    let pod_requests = total_pod_resources(pod);
    return pod_requests.cpu <= avail.cpu && pod_requests.memory <= avail.memory;
}

fn q(s: &str) -> ParsedQuantity {
    return s.try_into().expect("harness literal");
}

#[test]
fn golden_dump_gv1() {
    // tests/golden/gv1.json (BASELINE.json config C1: 10 pods x 5 nodes, resource_fits only)
    let nodes = vec![
        node("n0", Some(("4", "8589934592"))),
        node("n1", Some(("2", "4294967296"))),
        node("n2", Some(("8", "17179869184"))),
        node("n3", Some(("1", "1073741824"))),
        node("n4", None),
    ];
    let bound = vec![
        pod("b1", Some("n1"), vec![container(Some(&[("cpu", "500m"), ("memory", "1073741824")]))]),
        pod("b2a", Some("n2"), vec![container(Some(&[("cpu", "2"), ("memory", "4294967296")]))]),
        pod("b2b", Some("n2"), vec![
            container(Some(&[("cpu", "250m"), ("memory", "268435456")])),
            container(Some(&[("cpu", "250m"), ("memory", "268435456")])),
        ]),
        pod("b3", Some("n3"), vec![container(Some(&[("cpu", "1"), ("memory", "1073741824")]))]),
        pod("b4", Some("n4"), vec![container(Some(&[("cpu", "100m"), ("memory", "1")]))]),
        pod("stray", Some("not-a-node"), vec![container(Some(&[("cpu", "64"), ("memory", "1")]))]),
    ];
    let pods = vec![
        pod("p0", None, vec![]),
        pod("p1", None, vec![container(Some(&[("cpu", "500m"), ("memory", "1073741824")]))]),
        pod("p2", None, vec![container(Some(&[("cpu", "1500m"), ("memory", "3221225472")]))]),
        pod("p3", None, vec![container(Some(&[("cpu", "1501m"), ("memory", "1")]))]),
        pod("p4", None, vec![container(Some(&[("cpu", "100m"), ("memory", "3221225473")]))]),
        pod("p5", None, vec![container(Some(&[("cpu", "4"), ("memory", "8589934592")]))]),
        pod("p6", None, vec![container(Some(&[("cpu", "5500m"), ("memory", "12348030976")]))]),
        pod("p7", None, vec![
            container(Some(&[("cpu", "2"), ("memory", "4294967296")])),
            container(Some(&[("cpu", "3500m"), ("memory", "8053063680")])),
        ]),
        pod("p8", None, vec![container(Some(&[("cpu", "5501m"), ("memory", "0")]))]),
        pod("p9", None, vec![container(None)]),
    ];
    // expected totals as canonical quantity strings (millicores / bytes); equality is the crate's own PartialEq
    let exp_req_cpu = ["0m", "500m", "1500m", "1501m", "100m", "4000m", "5500m", "5500m", "5501m", "0m"];
    let exp_req_mem = ["0", "1073741824", "3221225472", "1", "3221225473", "8589934592", "12348030976", "12348030976", "0", "0"];
    let exp_free_cpu = ["4000m", "1500m", "5500m", "0m", "-100m"];
    let exp_free_mem = ["8589934592", "3221225472", "12348030976", "0", "-1"];

    let mut avail = Vec::new();
    for n in &nodes {
        let name = n.metadata.name.clone().unwrap();
        let on_node: Vec<&corev1::Pod> =
            bound.iter().filter(|p| p.spec.as_ref().unwrap().node_name.as_deref() == Some(name.as_str())).collect();
        avail.push(available(n, &on_node));
    }
    let mut free_ok = Vec::new();
    let mut free_show = Vec::new();
    for (i, a) in avail.iter().enumerate() {
        free_ok.push(a.cpu == q(exp_free_cpu[i]) && a.memory == q(exp_free_mem[i]));
        free_show.push(format!("[\"{}\",\"{}\"]", a.cpu, a.memory));
    }
    let mut req_ok = Vec::new();
    let mut req_show = Vec::new();
    let mut rows = Vec::new();
    for (i, p) in pods.iter().enumerate() {
        let r = total_pod_resources(p);
        req_ok.push(r.cpu == q(exp_req_cpu[i]) && r.memory == q(exp_req_mem[i]));
        req_show.push(format!("[\"{}\",\"{}\"]", r.cpu, r.memory));
        let row: String = avail.iter().map(|a| if fits(p, a) { '1' } else { '0' }).collect();
        rows.push(format!("\"{}\"", row));
    }
    println!(
        "GOLDEN {{\"kind\":\"gv1\",\"free_equals_expected\":{:?},\"free_display\":[{}],\"req_equals_expected\":{:?},\"req_display\":[{}],\"feasible_rows\":[{}]}}",
        free_ok,
        free_show.join(","),
        req_ok,
        req_show.join(","),
        rows.join(",")
    );
}

#[test]
fn golden_dump_quantities() {
    // tests/golden/quantity_kats.json "ok" entries: (string, value in 1/1000 units).  Inside the exact domain
    // (integer cores / millicores, plain integer bytes) the crate must agree with plain integer arithmetic; the
    // Ki/Mi/Gi and fractional entries show what the crate does OUTSIDE it (mixed-format normalisation).
    let kats: [(&str, &str); 14] = [
        ("0", "0m"), ("1", "1000m"), ("4", "4000m"), ("250m", "250m"), ("1500m", "1500m"), ("0m", "0m"), ("100", "100000m"),
        ("1073741824", "1073741824000m"), ("1Ki", "1024000m"), ("1Mi", "1048576000m"), ("1Gi", "1073741824000m"),
        ("1.5Gi", "1610612736000m"), ("1k", "1000000m"), ("0.5", "500m"),
    ];
    let mut out = Vec::new();
    for (s, milli) in kats.iter() {
        let parsed: Result<ParsedQuantity, _> = (*s).try_into();
        match parsed {
            Ok(v) => {
                // what total_pod_resources does with it: "0" += v  (src/util.rs:25-26,65)
                let mut acc: ParsedQuantity = "0".try_into().unwrap();
                acc += v.clone();
                let direct = v == q(milli);
                let accumulated = acc == q(milli);
                let le = acc <= q(milli) && q(milli) <= acc;
                out.push(format!("[\"{}\",\"{}\",{},{},{}]", s, acc, direct, accumulated, le));
            },
            Err(_) => out.push(format!("[\"{}\",null,false,false,false]", s)),
        }
    }
    println!("GOLDEN {{\"kind\":\"quantities\",\"rows\":[{}]}}", out.join(","));
}
