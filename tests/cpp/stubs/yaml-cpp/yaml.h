// Stand-in for <yaml-cpp/yaml.h> (yaml-cpp is not installed in this image): just enough of YAML::Node for the
// reference's constructors -- node["key"], node[index], node.as<T>() -- so that the drop-in classes compile and link
// against the REFERENCE'S OWN interface headers with -DB2_WITH_YAML (tests/test_dropin_reference_headers.py).
// TEST INFRASTRUCTURE ONLY: declarations + a trivial in-memory tree, no parser.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace YAML {
class Node {
  public:
    Node() : d_(std::make_shared<Data>()) {}
    template <typename T> explicit Node(const T& scalar) : d_(std::make_shared<Data>()) { std::ostringstream o; o << scalar; d_->scalar = o.str(); }
    Node operator[](const std::string& key) const { return d_->map[key]; }
    Node operator[](const char* key) const { return d_->map[std::string(key)]; }
    Node operator[](int i) const { return d_->seq.at((size_t)i); }
    Node operator[](size_t i) const { return d_->seq.at(i); }
    template <typename T> T as() const { std::istringstream in(d_->scalar); T v{}; in >> v; return v; }
    // builders used by the test only
    Node& set(const std::string& key, const Node& v) { d_->map[key] = v; return *this; }
    Node& push(const Node& v) { d_->seq.push_back(v); return *this; }

  private:
    struct Data { std::string scalar; std::map<std::string, Node> map; std::vector<Node> seq; };
    std::shared_ptr<Data> d_;
};
template <> inline std::string Node::as<std::string>() const { return d_->scalar; }
}  // namespace YAML
