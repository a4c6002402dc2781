# Optional public bootstrap peer: a tiny CPU node with a stable address that new nodes dial first.
variable "bootstrap_enabled" { default = false }

resource "aws_eip" "bootstrap" {
  count  = var.bootstrap_enabled ? 1 : 0
  domain = "vpc"
}
resource "aws_instance" "bootstrap" {
  count                  = var.bootstrap_enabled ? 1 : 0
  ami                    = data.aws_ami.dl.id
  instance_type          = "t3.small"
  subnet_id              = aws_subnet.dmz.id
  vpc_security_group_ids = [aws_security_group.crawler.id]
  user_data = <<-EOT
    #!/bin/bash
    mkdir -p /data && chown 1000:1000 /data
    ${local.run} -p 4001:4001 -e INFOMESH_NODE_ROLE=crawler -e INFOMESH_GPU_ENABLED=false \
      -e INFOMESH_RESOURCES_PROFILE=minimal ${var.image} _serve --no-crawl
  EOT
  tags = { Name = "infomesh-bootstrap" }
}
resource "aws_eip_association" "bootstrap" {
  count         = var.bootstrap_enabled ? 1 : 0
  instance_id   = aws_instance.bootstrap[0].id
  allocation_id = aws_eip.bootstrap[0].id
}
output "bootstrap_multiaddr" {
  value = var.bootstrap_enabled ? "/ip4/${aws_eip.bootstrap[0].public_ip}/tcp/4001" : ""
}
