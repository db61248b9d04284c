// fastparse.hpp — snapshot ingest without a DOM: v1.Pod / v1.Node items of a LIST are scanned once and the few
// fields the hot path reads go straight into the object model (objects.hpp); everything else is skipped by a string-aware
// bracket scan. What SyncWithClient copies per object (pkg/framework/simulator.go:176-295) is the whole API object; what the
// scheduler plugins of this path read of an EXISTING pod is: phase, nodeName, namespace, labels, deletionTimestamp, the
// containers' requests and hostPorts, init containers / overhead / pod-level resources, container statuses (in-place resize),
// pod (anti-)affinity terms and priority; of a node: name, labels, taints, unschedulable, allocatable, images.
//
// The fast path handles the common shape and gives up (returns false: the caller parses the item with the DOM parser and
// Pod::parse / Node::parse) as soon as it meets a construct whose handling needs the general code: escaped strings, pod
// affinity, init containers, overhead, pod-level resources, resize conditions, non-string label values. It never guesses.
#pragma once
#include <cstring>
#include <string_view>
#include "objects.hpp"

namespace cch {
namespace fast {

struct Bail {};   // thrown to leave the fast path (malformed input is left to the DOM parser, which reports it properly)

struct Cur {
  const char *p, *e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  char peek() { ws(); if (p >= e) throw Bail(); return *p; }
  void expect(char c) { if (peek() != c) throw Bail(); p++; }
  bool maybe(char c) { if (peek() == c) { p++; return true; } return false; }
  // raw contents of a string without escapes (an escape sequence leaves the fast path)
  std::string_view str() {
    expect('"');
    const char *b = p;
    const char *q = (const char *)memchr(p, '"', (size_t)(e - p));
    if (!q) throw Bail();
    if (memchr(b, '\\', (size_t)(q - b))) throw Bail();
    p = q + 1;
    return std::string_view(b, (size_t)(q - b));
  }
  void skip_string() {
    expect('"');
    for (; p < e; p++) { if (*p == '\\') p++; else if (*p == '"') { p++; return; } }
    throw Bail();
  }
  void skip() {   // any value
    const char c = peek();
    if (c == '"') { skip_string(); return; }
    if (c == '{' || c == '[') {
      int depth = 0;
      for (; p < e; p++) {
        const char d = *p;
        if (d == '"') { skip_string(); p--; continue; }
        if (d == '{' || d == '[') depth++;
        else if (d == '}' || d == ']') { if (--depth == 0) { p++; return; } }
      }
      throw Bail();
    }
    while (p < e && *p != ',' && *p != ']' && *p != '}' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
  }
  bool is_null() { ws(); return e - p >= 4 && memcmp(p, "null", 4) == 0; }
  long long integer() {
    ws();
    const char *b = p;
    if (p < e && (*p == '-' || *p == '+')) p++;
    while (p < e && *p >= '0' && *p <= '9') p++;
    if (p == b || (p < e && (*p == '.' || *p == 'e' || *p == 'E'))) throw Bail();
    return strtoll(std::string(b, (size_t)(p - b)).c_str(), nullptr, 10);
  }
  bool boolean() {
    ws();
    if (e - p >= 4 && memcmp(p, "true", 4) == 0) { p += 4; return true; }
    if (e - p >= 5 && memcmp(p, "false", 5) == 0) { p += 5; return false; }
    throw Bail();
  }
  // a Quantity: a string ("100m") or a bare number (cpu: 2)
  Quantity quantity() {
    if (peek() == '"') return Quantity::parse(std::string(str()));
    const char *b = p;
    while (p < e && *p != ',' && *p != ']' && *p != '}' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
    if (p == b) throw Bail();
    return Quantity::parse(std::string(b, (size_t)(p - b)));
  }
};

// object members: f(key) must consume the value
template <class F> inline void members(Cur &c, F f) {
  if (c.is_null()) { c.p += 4; return; }
  c.expect('{');
  if (c.maybe('}')) return;
  for (;;) {
    const std::string_view k = c.str();
    c.expect(':');
    f(k);
    if (c.maybe(',')) continue;
    c.expect('}');
    return;
  }
}
template <class F> inline void elements(Cur &c, F f) {
  if (c.is_null()) { c.p += 4; return; }
  c.expect('[');
  if (c.maybe(']')) return;
  for (;;) {
    f();
    if (c.maybe(',')) continue;
    c.expect(']');
    return;
  }
}

inline void string_map(Cur &c, Labels &out) {
  members(c, [&](std::string_view k) { if (c.peek() != '"') throw Bail(); out[std::string(k)] = std::string(c.str()); });
}
inline void resource_list(Cur &c, ResourceList &out) {
  members(c, [&](std::string_view k) { out[std::string(k)] = c.quantity(); });
}

inline void container(Cur &c, Container &k) {
  members(c, [&](std::string_view key) {
    if (key == "name") k.name = std::string(c.str());
    else if (key == "image") k.image = std::string(c.str());
    else if (key == "restartPolicy") k.restart_always = c.str() == "Always";
    else if (key == "resources") members(c, [&](std::string_view r) { if (r == "requests") resource_list(c, k.requests); else c.skip(); });
    else if (key == "ports")
      elements(c, [&] {
        ContainerPort cp;
        members(c, [&](std::string_view pk) {
          if (pk == "hostPort") cp.host_port = (int)c.integer();
          else if (pk == "hostIP") cp.host_ip = std::string(c.str());
          else if (pk == "protocol") cp.protocol = std::string(c.str());
          else c.skip();
        });
        k.ports.push_back(cp);
      });
    else c.skip();
  });
}

// An existing pod of the snapshot. Returns false when the item needs the general parser.
inline bool pod(std::string_view item, Pod &p) {
  try {
    Cur c{item.data(), item.data() + item.size()};
    bool have_owner = false;
    members(c, [&](std::string_view top) {
      if (top == "metadata")
        members(c, [&](std::string_view k) {
          if (k == "name") p.name = std::string(c.str());
          else if (k == "namespace") p.ns = std::string(c.str());
          else if (k == "labels") string_map(c, p.labels);
          else if (k == "deletionTimestamp") { p.terminating = !c.is_null(); c.skip(); }
          else if (k == "ownerReferences")
            elements(c, [&] {
              std::string av, kind, name; bool ctrl = false;
              members(c, [&](std::string_view ok) {
                if (ok == "apiVersion") av = std::string(c.str());
                else if (ok == "kind") kind = std::string(c.str());
                else if (ok == "name") name = std::string(c.str());
                else if (ok == "controller") { if (c.is_null()) c.skip(); else ctrl = c.boolean(); }
                else c.skip();
              });
              if (ctrl && !have_owner) { have_owner = true; p.owner_api_version = av; p.owner_kind = kind; p.owner_name = name; }
            });
          else c.skip();
        });
      else if (top == "spec")
        members(c, [&](std::string_view k) {
          if (k == "nodeName") p.node_name = std::string(c.str());
          else if (k == "schedulerName") p.scheduler_name = std::string(c.str());
          else if (k == "preemptionPolicy") p.preemption_policy = std::string(c.str());
          else if (k == "priority") { if (c.is_null()) c.skip(); else p.priority = (int)c.integer(); }
          else if (k == "containers") elements(c, [&] { p.containers.emplace_back(); container(c, p.containers.back()); });
          else if (k == "nodeSelector") { p.has_node_selector = !c.is_null(); string_map(c, p.node_selector); }
          // the general parser's business: init-container / overhead / pod-level accounting rules, pod (anti-)affinity terms.
          // (tolerations, topologySpreadConstraints, nodeAffinity, volumes, resourceClaims and schedulingGates of an EXISTING pod are
          //  read by nobody on this path — they only matter on the simulated pod, which always takes the general parser.)
          else if (k == "initContainers") { if (c.is_null()) c.skip(); else { c.expect('['); if (!c.maybe(']')) throw Bail(); } }
          else if (k == "overhead" || k == "resources") { if (c.is_null()) c.skip(); else throw Bail(); }
          else if (k == "affinity")
            members(c, [&](std::string_view ak) {
              if ((ak == "podAffinity" || ak == "podAntiAffinity") && !c.is_null()) throw Bail();
              c.skip();
            });
          else c.skip();
        });
      else if (top == "status")
        members(c, [&](std::string_view k) {
          if (k == "phase") p.phase = std::string(c.str());
          else if (k == "containerStatuses" || k == "initContainerStatuses")
            elements(c, [&] {
              std::string name; ResourceList req, alloc; bool has_res = false, has_alloc = false;
              members(c, [&](std::string_view sk) {
                if (sk == "name") name = std::string(c.str());
                else if (sk == "resources") { has_res = !c.is_null(); members(c, [&](std::string_view r) { if (r == "requests") resource_list(c, req); else c.skip(); }); }
                else if (sk == "allocatedResources") { has_alloc = !c.is_null(); resource_list(c, alloc); }
                else c.skip();
              });
              if (has_res) p.status_resources[name] = req;
              if (has_alloc) p.status_allocated[name] = alloc;
            });
          else if (k == "conditions")
            elements(c, [&] {
              std::string type, reason;
              members(c, [&](std::string_view ck) {
                if (ck == "type") type = std::string(c.str());
                else if (ck == "reason") reason = std::string(c.str());
                else c.skip();
              });
              if (type == "PodResizePending") p.resize_infeasible = reason == "Infeasible";
            });
          else c.skip();
        });
      else c.skip();
    });
    c.ws();
    if (c.p != c.e) return false;
    if (p.ns.empty()) p.ns = "default";
    return true;
  } catch (const Bail &) { return false; }
}

inline bool node(std::string_view item, Node &n) {
  try {
    Cur c{item.data(), item.data() + item.size()};
    members(c, [&](std::string_view top) {
      if (top == "metadata")
        members(c, [&](std::string_view k) {
          if (k == "name") n.name = std::string(c.str());
          else if (k == "labels") string_map(c, n.labels);
          else c.skip();
        });
      else if (top == "spec")
        members(c, [&](std::string_view k) {
          if (k == "unschedulable") { if (c.is_null()) c.skip(); else n.unschedulable = c.boolean(); }
          else if (k == "taints")
            elements(c, [&] {
              Taint t;
              members(c, [&](std::string_view tk) {
                if (tk == "key") t.key = std::string(c.str());
                else if (tk == "value") t.value = std::string(c.str());
                else if (tk == "effect") t.effect = std::string(c.str());
                else c.skip();
              });
              n.taints.push_back(t);
            });
          else c.skip();
        });
      else if (top == "status")
        members(c, [&](std::string_view k) {
          if (k == "allocatable") resource_list(c, n.allocatable);
          else if (k == "images")
            elements(c, [&] {
              std::vector<std::string> names; long long size = 0;
              members(c, [&](std::string_view ik) {
                if (ik == "names") elements(c, [&] { names.push_back(std::string(c.str())); });
                else if (ik == "sizeBytes") { if (c.is_null()) c.skip(); else size = c.integer(); }
                else c.skip();
              });
              for (auto &nm : names) { n.image_names.push_back(nm); n.images.push_back({nm, size}); }
            });
          else c.skip();
        });
      else c.skip();
    });
    c.ws();
    return c.p == c.e;
  } catch (const Bail &) { return false; }
}

}  // namespace fast
}  // namespace cch
